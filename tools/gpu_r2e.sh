#!/bin/bash
mkdir -p gpurun_out
T=r2e
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/${T}_pytest.log
for V in 1 0; do
SCORE_COATT_LEAN=$V timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${T}_lean$V.json 2> gpurun_out/${T}_lean$V.err; echo "lean=$V rc=$?"
python tools/bench_digest.py gpurun_out/${T}_lean$V.json 2>/dev/null || cut -c1-400 gpurun_out/${T}_lean$V.json
done
for V in 1 0; do
SCORE_COATT_LEAN=$V timeout 300 python bench.py --steps 50 --warmup 5 --batch 16384 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${T}_b16k_lean$V.json 2> gpurun_out/${T}_b16k_lean$V.err; echo "b16k lean=$V rc=$?"
python tools/bench_digest.py gpurun_out/${T}_b16k_lean$V.json 2>/dev/null || cut -c1-400 gpurun_out/${T}_b16k_lean$V.json
done
