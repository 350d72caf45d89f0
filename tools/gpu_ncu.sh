#!/bin/bash
# ncu captures (one GPU).  Usage: tools/gpu_ncu.sh <tag> [extra bench args]
TAG=${1:-r01}; shift
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 $@"
# the HBM-side kernels, full set with source
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'coatt_fwd_kernel|coatt_bwd_kernel|emb_update_kernel|emb_catchup_rows_kernel' -s 16 -c 4 \
  -o gpurun_out/${TAG}_emb -f $B > gpurun_out/${TAG}_ncu_emb.log 2>&1
echo "emb rc=$?"
# one whole step, every kernel, light sections (the 64 MiB return limit)
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --section WarpStateStats --section SchedulerStats --clock-control none -s 400 -c 50 \
  -o gpurun_out/${TAG}_step -f $B > gpurun_out/${TAG}_ncu_step.log 2>&1
echo "step rc=$?"
du -sh gpurun_out; ls -la gpurun_out/
