#!/bin/bash
# warm launch list + full captures of selected kernels.  Usage: tools/gpu_ncu2.sh <tag> <kernel-regex> [bench args]
TAG=$1; RX=$2; shift; shift
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 $@"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 300 --csv \
  --log-file gpurun_out/${TAG}_launches_warm.csv $B > gpurun_out/${TAG}_ncu_warm.log 2>&1
echo "warm rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$RX" -s 12 -c 6 \
  -o gpurun_out/${TAG}_sel -f $B > gpurun_out/${TAG}_ncu_sel.log 2>&1
echo "sel rc=$?"
du -sh gpurun_out
