#!/bin/bash
# One GPU-box pass: parity tests, bench line, optional ncu launch list.  Usage: tools/gpu_check.sh <tag> [ncu] [bench args...]
TAG=${1:-r01}; shift
NCU=0; if [ "$1" = "ncu" ]; then NCU=1; shift; fi
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 200 --warmup 10 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
if [ $NCU = 1 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu rc=$?"
fi
