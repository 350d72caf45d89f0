#!/bin/bash
# 8-GPU pass: peer-memory exchange trial at 8, Taobao dp8 with / without it, Taobao row-sharded, config 4 (CCMR K=20 dp) at 4 and 8,
# then the default bench line (probe + large-vocab leg)
mkdir -p gpurun_out
T=r2r
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node"
short() { python tools/bench_digest.py "$1" 2>/dev/null || cut -c1-300 "$1"; }
SCORE_DP_P2P=1 timeout 200 $TR 8 tools/multigpu_check.py tiny_tb > gpurun_out/${T}_multi_check_p2p8.log 2>&1; echo "p2p check rc=$?"; grep -h "MULTIGPU_CHECK" gpurun_out/${T}_multi_check_p2p8.log
SCORE_DP_P2P=1 timeout 200 $TR 8 bench.py --gpus 8 --parallel dp --no-large-vocab --steps 100 --warmup 10 > gpurun_out/${T}_tb_dp8_p2p.json 2> gpurun_out/${T}_tb_dp8_p2p.err; echo "dp8 p2p rc=$?"; short gpurun_out/${T}_tb_dp8_p2p.json
timeout 200 $TR 8 bench.py --gpus 8 --parallel dp --no-large-vocab --steps 100 --warmup 10 > gpurun_out/${T}_tb_dp8.json 2> gpurun_out/${T}_tb_dp8.err; echo "dp8 rc=$?"; short gpurun_out/${T}_tb_dp8.json
timeout 200 $TR 8 bench.py --gpus 8 --parallel sharded --no-large-vocab --steps 100 --warmup 10 > gpurun_out/${T}_tb_sh8.json 2> gpurun_out/${T}_tb_sh8.err; echo "sharded8 rc=$?"; short gpurun_out/${T}_tb_sh8.json
for N in 4 8; do
timeout 300 $TR $N bench.py --gpus $N --workload ccmr_k20 --parallel dp --steps 30 --warmup 5 --e2e-steps 5 > gpurun_out/${T}_ccmrk20_dp$N.json 2> gpurun_out/${T}_ccmrk20_dp$N.err; echo "ccmr_k20 dp$N rc=$?"; short gpurun_out/${T}_ccmrk20_dp$N.json
done
timeout 600 $TR 8 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/${T}_bench8.json 2> gpurun_out/${T}_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2r_bench8.json').read().strip().splitlines()[-1])
    print("value %.0f ms %.3f e2e %.0f loss %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss']), (d.get('parallelism_scheme') or {}).get('probe_ms_per_step'))
    lv=d.get('large_vocab'); print("large_vocab: %.0f samples/s %.3f ms, x %.2f (1 GPU %.3f ms)"%(lv['value'],lv['ms_per_step'],lv['x_vs_1gpu_shard'],lv['one_gpu_shard']['ms_per_step']))
except Exception as e: print("no json",e)
PY
timeout 300 $TR 8 tools/shard_timeline.py large_vocab 30 2>/dev/null | grep SHARD_TIMELINE | tee gpurun_out/${T}_shard_timeline_lv_8.txt
