#!/bin/bash
mkdir -p gpurun_out
T=r2m
for W in ccmr ccmr_k20; do
timeout 600 python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${T}_$W.json 2> gpurun_out/${T}_$W.err; echo "$W rc=$?"
python tools/bench_digest.py gpurun_out/${T}_$W.json
done
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${T}_taobao.json 2> gpurun_out/${T}_taobao.err; echo "taobao rc=$?"
python tools/bench_digest.py gpurun_out/${T}_taobao.json
