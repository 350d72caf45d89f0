#!/bin/bash
# N-GPU: config 4 (CCMR K=20, dense data-parallel) bench line; optional peer-memory exchange trial (TRY_P2P=1)
mkdir -p gpurun_out
T=r2p
N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --workload ccmr_k20 --parallel dp --steps 30 --warmup 5 --e2e-steps 5 > gpurun_out/${T}_ccmrk20_dp$N.json 2> gpurun_out/${T}_ccmrk20_dp$N.err; echo "ccmr_k20 dp$N rc=$?"
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${T}_ccmrk20_dp$N.err | tail -3
python tools/bench_digest.py gpurun_out/${T}_ccmrk20_dp$N.json
if [ "$TRY_P2P" = "1" ]; then
SCORE_DP_P2P=1 timeout 150 $RUN tools/multigpu_check.py tiny_tb > gpurun_out/${T}_multi_check_p2p.log 2>&1; echo "p2p check rc=$?"; grep -h "MULTIGPU_CHECK\|divergence\|Error\|error" gpurun_out/${T}_multi_check_p2p.log | head -5
SCORE_DP_P2P=1 timeout 200 $RUN bench.py --gpus $N --parallel dp --no-large-vocab --steps 100 --warmup 10 > gpurun_out/${T}_bench_dp${N}_p2p.json 2> gpurun_out/${T}_bench_dp${N}_p2p.err; echo "p2p dp bench rc=$?"
python tools/bench_digest.py gpurun_out/${T}_bench_dp${N}_p2p.json
timeout 200 $RUN bench.py --gpus $N --parallel dp --no-large-vocab --steps 100 --warmup 10 > gpurun_out/${T}_bench_dp${N}.json 2> gpurun_out/${T}_bench_dp${N}.err; echo "dp bench rc=$?"
python tools/bench_digest.py gpurun_out/${T}_bench_dp${N}.json
fi
