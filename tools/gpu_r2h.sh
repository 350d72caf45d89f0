#!/bin/bash
# 8-GPU pass: bench line (Taobao dp8) + large-vocab leg
mkdir -p gpurun_out
T=r2h
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/${T}_bench8.json 2> gpurun_out/${T}_bench8.err; echo "bench8 rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${T}_bench8.err | tail -5
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench8.json').read().strip().splitlines()[-1])
    print("value %.0f ms %.3f e2e %.0f loss %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss']))
    print(json.dumps(d.get('large_vocab'),indent=1))
except Exception as e: print("no json",e)
PY
