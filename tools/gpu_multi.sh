#!/bin/bash
# N-GPU pass (gpurun --gpus N): split-step parity vs 1 GPU, phase timeline, then bench lines.  Usage: tools/gpu_multi.sh <tag> <N> [nosharded]
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $RUN tools/multigpu_check.py tiny_tb > gpurun_out/${TAG}_multi_check.log 2>&1; echo "check rc=$?"; grep -h "MULTIGPU_CHECK\|divergence" gpurun_out/${TAG}_multi_check.log
timeout 300 $RUN tools/dp_timeline.py 2>/dev/null | grep DP_TIMELINE | tee gpurun_out/${TAG}_dp_timeline.txt
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_dp$N.json 2> gpurun_out/${TAG}_bench_dp$N.err; echo "dp bench rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_dp$N.err | tail -3; cut -c1-330 gpurun_out/${TAG}_bench_dp$N.json
if [ "$3" != "nosharded" ]; then
timeout 600 $RUN bench.py --gpus $N --steps 30 --warmup 5 --workload large_vocab_shard --parallel sharded > gpurun_out/${TAG}_bench_sh$N.json 2> gpurun_out/${TAG}_bench_sh$N.err; echo "sharded bench rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_sh$N.err | tail -3; cut -c1-330 gpurun_out/${TAG}_bench_sh$N.json
fi
if [ "$TRY_P2P" = "1" ]; then
# opt-in peer-memory exchange (SCORE_DP_P2P=1, parallel.py): parity first, then the bench line; short timeouts - a
# missing peer signal would hang the barrier
SCORE_DP_P2P=1 timeout 120 $RUN tools/multigpu_check.py tiny_tb > gpurun_out/${TAG}_multi_check_p2p.log 2>&1; echo "p2p check rc=$?"; grep -h "MULTIGPU_CHECK\|divergence\|Error\|error" gpurun_out/${TAG}_multi_check_p2p.log | head -5
SCORE_DP_P2P=1 timeout 120 $RUN tools/dp_timeline.py 2>/dev/null | grep DP_TIMELINE | tee gpurun_out/${TAG}_dp_timeline_p2p.txt
SCORE_DP_P2P=1 timeout 200 $RUN bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_dp${N}_p2p.json 2> gpurun_out/${TAG}_bench_dp${N}_p2p.err; echo "p2p dp bench rc=$?"
cut -c1-330 gpurun_out/${TAG}_bench_dp${N}_p2p.json
fi
