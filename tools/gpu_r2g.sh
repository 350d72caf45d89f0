#!/bin/bash
# 2-GPU pass: GPU tests (incl. real-NCCL dp + sharded parity), then the bench line with the large-vocab leg
mkdir -p gpurun_out
T=r2g
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/${T}_pytest.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; echo "bench2 rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${T}_bench2.err | tail -5
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2g_bench2.json').read().strip().splitlines()[-1])
    print("value %.0f ms %.3f e2e %.0f loss %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss']))
    print(json.dumps(d.get('large_vocab'),indent=1))
except Exception as e: print("no json",e)
PY
