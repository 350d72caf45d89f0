#!/bin/bash
mkdir -p gpurun_out
T=r2j
N=${1:-2}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${T}_pytest.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $RUN bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err; echo "bench rc=$?"
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${T}_bench$N.err | tail -5
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 10 --parallel sharded --no-large-vocab > gpurun_out/${T}_bench${N}_tbsh.json 2> gpurun_out/${T}_bench${N}_tbsh.err; echo "taobao sharded rc=$?"
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${T}_bench${N}_tbsh.err | tail -5
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ('gpurun_out/r2j_bench%s.json'%N,'gpurun_out/r2j_bench%s_tbsh.json'%N):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.3f e2e %.0f loss %s"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['final_loss']), d['config']['parallelism'][:40])
        lv=d.get('large_vocab')
        if lv: print("  large_vocab: %.0f samples/s, %.3f ms, x_vs_1gpu_shard %.2f (1 GPU %.3f ms)"%(lv['value'],lv['ms_per_step'],lv['x_vs_1gpu_shard'],lv['one_gpu_shard']['ms_per_step']))
    except Exception as e: print(f,"no json",e)
PY
