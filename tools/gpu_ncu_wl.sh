#!/bin/bash
# full ncu capture of the HBM-side kernels of one workload at a late step.  Usage: tools/gpu_ncu_wl.sh <tag> <workload> [skip] [count]
TAG=$1; W=$2; SKIP=${3:-8}; CNT=${4:-4}
mkdir -p gpurun_out
B="python bench.py --workload $W --steps 6 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'coatt_fwd_kernel|coatt_bwd_kernel|emb_update_kernel|emb_replay_kernel|build_keys_kernel' -s $SKIP -c $CNT \
  -o gpurun_out/${TAG}_${W} -f $B > gpurun_out/${TAG}_${W}_ncu.log 2>&1
echo "$W ncu rc=$?"; tail -2 gpurun_out/${TAG}_${W}_ncu.log
