#!/bin/bash
# second A/B pass: split sort (SCORE_SCHED bit 4) and ring-depth variants around the new default.  Usage: tools/gpu_ab2.sh <tag>
TAG=${1:-ab2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
short() { python tools/bench_digest.py "$1" 2>/dev/null || cut -c1-300 "$1"; }
for S in 7 3 7 3; do
  SCORE_SCHED=$S timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_sched$S.json 2> gpurun_out/${TAG}_sched$S.err
  echo "SCHED=$S rc=$? $(short gpurun_out/${TAG}_sched$S.json)"
done
for V in kc16_nst6 kc16_nst12 kc16_nst4; do
  SCORE_B200_LIB=$PWD/score_b200/libscore_b200.$V.so timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  echo "VARIANT=$V rc=$? $(short gpurun_out/${TAG}_$V.json)"
done
for W in tmall ccmr large_vocab_shard; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 5 --workload $W > gpurun_out/${TAG}_wl_$W.json 2> gpurun_out/${TAG}_wl_$W.err
  echo "WORKLOAD=$W rc=$? $(short gpurun_out/${TAG}_wl_$W.json)"
done
