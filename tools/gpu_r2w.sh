#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node"
SCORE_SHARD_P2P=1 timeout 45 $TR 8 tools/shard_timeline.py large_vocab 20 2>gpurun_out/r2w_tl.err | grep SHARD_TIMELINE | tee gpurun_out/r2w_shard_timeline_lv_p2p_8.txt
grep -i "error\|Traceback" gpurun_out/r2w_tl.err | head -3
