#!/bin/bash
# N-GPU pass (gpurun --gpus N): split-step parity vs 1 GPU, then bench lines.  Usage: tools/gpu_multi.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN tools/multigpu_check.py tiny_tb > gpurun_out/${TAG}_multi_check.log 2>&1; echo "check rc=$?"; tail -12 gpurun_out/${TAG}_multi_check.log
timeout 600 $RUN bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_dp$N.json 2> gpurun_out/${TAG}_bench_dp$N.err; echo "dp bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_dp$N.err; cat gpurun_out/${TAG}_bench_dp$N.json | cut -c1-600
timeout 600 $RUN bench.py --gpus $N --steps 30 --warmup 5 --workload large_vocab_shard --parallel sharded > gpurun_out/${TAG}_bench_sh$N.json 2> gpurun_out/${TAG}_bench_sh$N.err; echo "sharded bench rc=$?"
tail -3 gpurun_out/${TAG}_bench_sh$N.err; cat gpurun_out/${TAG}_bench_sh$N.json | cut -c1-600
