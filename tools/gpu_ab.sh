#!/bin/bash
# One GPU-box pass for A/B decisions: full parity tests, the default bench line, then short bench runs over the
# scheduling knob (SCORE_SCHED bits: 1 att_qb on the side stream, 2 dense reduce next to the embedding update,
# 4 sort forked after the gather) and the compile-time tile-ring variants.  Usage: tools/gpu_ab.sh <tag>
TAG=${1:-ab}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
short() { python tools/bench_digest.py "$1" 2>/dev/null || cut -c1-300 "$1"; }
for S in 0 1 2 4 3 7; do
  SCORE_SCHED=$S timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_sched$S.json 2> gpurun_out/${TAG}_sched$S.err
  echo "SCHED=$S rc=$? $(short gpurun_out/${TAG}_sched$S.json)"
done
for V in kc16_nst8 kc32_nst4 kc32_nst6; do
  SCORE_B200_LIB=$PWD/score_b200/libscore_b200.$V.so timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  echo "VARIANT=$V rc=$? $(short gpurun_out/${TAG}_$V.json)"
done
