#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN tools/shard_timeline.py large_vocab 30 2>/dev/null | grep SHARD_TIMELINE | tee gpurun_out/r2k_shard_timeline_lv_$N.txt
timeout 600 $RUN tools/shard_timeline.py taobao 50 2>/dev/null | grep SHARD_TIMELINE | tee gpurun_out/r2k_shard_timeline_tb_$N.txt
