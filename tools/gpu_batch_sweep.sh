#!/bin/bash
# kernel rooflines of one workload at growing batch sizes.  Usage: tools/gpu_batch_sweep.sh <tag> <workload> <batches...>
TAG=$1; W=$2; shift; shift
mkdir -p gpurun_out
for B in "$@"; do
  timeout 600 python bench.py --workload $W --batch $B --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_${W}_b$B.json 2> gpurun_out/${TAG}_${W}_b$B.err
  echo "$W B=$B rc=$?"; tail -2 gpurun_out/${TAG}_${W}_b$B.err
done
python tools/bench_digest.py gpurun_out/${TAG}_${W}_b*.json
