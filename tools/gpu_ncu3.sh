#!/bin/bash
# full captures of selected kernels at a LATE step (steady state of the lazy optimizer).  Usage: tools/gpu_ncu3.sh <tag> <regex> <skip> <count>
TAG=$1; RX=$2; SKIP=$3; CNT=$4
mkdir -p gpurun_out
B="python bench.py --steps 120 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
  -o gpurun_out/${TAG}_late -f $B > gpurun_out/${TAG}_ncu_late.log 2>&1
echo "late rc=$?"; tail -3 gpurun_out/${TAG}_ncu_late.log
du -sh gpurun_out
