#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node"
SCORE_BENCH_WATCHDOG=110 timeout 170 $TR $N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2t_bench$N.json 2> gpurun_out/r2t_bench$N.err; echo "bench rc=$?"
grep -n "File \"/root\|File \"/tmp\|bench.py\|parallel.py\|Thread\|most recent" gpurun_out/r2t_bench$N.err | head -40
cut -c1-300 gpurun_out/r2t_bench$N.json
