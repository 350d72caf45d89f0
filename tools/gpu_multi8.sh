#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): parity check + the two scaling bench lines.  Usage: tools/gpu_multi8.sh <tag> [N]
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $RUN tools/multigpu_check.py tiny_tb > gpurun_out/${TAG}_multi_check$N.log 2>&1; echo "check rc=$?"; grep -E "MULTIGPU|replica" gpurun_out/${TAG}_multi_check$N.log | tail -3
timeout 400 $RUN bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_dp$N.json 2> gpurun_out/${TAG}_bench_dp$N.err; echo "dp bench rc=$?"
grep -v "^\*\|OMP" gpurun_out/${TAG}_bench_dp$N.err | tail -3; cut -c1-200 gpurun_out/${TAG}_bench_dp$N.json
timeout 600 $RUN bench.py --gpus $N --steps 30 --warmup 5 --workload large_vocab --parallel sharded > gpurun_out/${TAG}_bench_lv$N.json 2> gpurun_out/${TAG}_bench_lv$N.err; echo "sharded bench rc=$?"
grep -v "^\*\|OMP" gpurun_out/${TAG}_bench_lv$N.err | tail -3; cut -c1-200 gpurun_out/${TAG}_bench_lv$N.json
