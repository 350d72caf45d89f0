#!/bin/bash
# short profile pass: launch list of one step + full ncu capture of the HBM-side kernels.  Usage: tools/gpu_prof2.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'coatt_fwd_kernel|coatt_bwd_kernel|emb_update_kernel|emb_replay_kernel|build_keys_kernel' -s 35 -c 5 \
  -o gpurun_out/${TAG}_full -f $B > gpurun_out/${TAG}_full_ncu.log 2>&1
echo "full rc=$?"; ls -la gpurun_out/${TAG}_full.ncu-rep
