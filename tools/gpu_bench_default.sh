#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node"
SCORE_BENCH_WATCHDOG=150 timeout 200 $TR $N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2u_bench$N.json 2> gpurun_out/r2u_bench$N.err; echo "bench rc=$?"
grep -n "File \"/root\|Thread\|Error" gpurun_out/r2u_bench$N.err | head -20
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2u_bench%s.json'%N).read().strip().splitlines()[-1])
    print("value %.0f ms %.3f e2e %.0f loss %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss']), (d.get('parallelism_scheme') or {}).get('probe_ms_per_step'), d['clocks'])
    lv=d.get('large_vocab'); print("large_vocab: %.0f samples/s %.3f ms, x %.2f (1 GPU %.3f ms)"%(lv['value'],lv['ms_per_step'],lv['x_vs_1gpu_shard'],lv['one_gpu_shard']['ms_per_step']))
except Exception as e: print("no json",e)
PY
