#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node"
for V in plain graph_dropout; do
SCORE_SHARD_P2P=1 timeout 120 $TR $N tools/multigpu_check.py tiny_tb $V > gpurun_out/r2v_check_p2p_$V.log 2>&1; echo "p2p check $V rc=$?"; grep -h "MULTIGPU_CHECK\|Error\|error" gpurun_out/r2v_check_p2p_$V.log | head -4
done
SCORE_SHARD_P2P=1 timeout 120 $TR $N tools/shard_timeline.py large_vocab 30 2>gpurun_out/r2v_tl.err | grep SHARD_TIMELINE | tee gpurun_out/r2v_shard_timeline_lv_p2p_$N.txt
timeout 120 $TR $N tools/shard_timeline.py large_vocab 30 2>/dev/null | grep SHARD_TIMELINE | tee gpurun_out/r2v_shard_timeline_lv_$N.txt
tail -3 gpurun_out/r2v_tl.err | cut -c1-200
