#!/bin/bash
# bench lines of the other BASELINE shapes on one GPU (kernel rooflines at scale).  Usage: tools/gpu_workloads.sh <tag> [workloads...]
TAG=$1; shift
mkdir -p gpurun_out
for W in "$@"; do
  timeout 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  echo "$W rc=$?"; tail -2 gpurun_out/${TAG}_bench_$W.err; cat gpurun_out/${TAG}_bench_$W.json
done
