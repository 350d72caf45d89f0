#!/bin/bash
# ncu evidence of round 2: full-set capture of the HBM-side kernels (lean gather / backward, scatter + Adam) in the 6th
# step of a Taobao run, and the launch list of one whole step
mkdir -p gpurun_out
T=r2n
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'coatt_fwd_lean_kernel|coatt_bwd_lean_kernel|emb_update_kernel' -s 15 -c 3 \
  -o gpurun_out/${T}_emb -f $B > gpurun_out/${T}_ncu_emb.log 2>&1
echo "emb rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
  --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/${T}_*
