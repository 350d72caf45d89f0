#!/bin/bash
# ncu --set full with source counters for the dense-middle kernels of one Taobao step
mkdir -p gpurun_out
T=r2s
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'fc_fwd_kernel|fc_bwd_kernel|att_fwd2_kernel|att_bwd2_kernel|gru_fwd_kernel|gru_bwd_kernel|rowgemm_kernel' -s 35 -c 8 \
  -o gpurun_out/${T}_dense -f $B > gpurun_out/${T}_ncu_dense.log 2>&1
echo "dense rc=$?"; ls -la gpurun_out/${T}_dense.ncu-rep
