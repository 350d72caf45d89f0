#!/bin/bash
# profile pass of the current state (one GPU): launch list of one step + full ncu capture of the main kernels.
# Usage: tools/gpu_prof.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1; head -40 gpurun_out/${TAG}_launches_summary.txt
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'coatt_fwd_kernel|coatt_bwd_kernel|emb_update_kernel|emb_replay_kernel|build_keys_kernel|fc_fwd_kernel|fc_bwd_kernel|att_fwd2_kernel|att_bwd2_kernel|gru_fwd_kernel|gru_bwd_kernel|rowgemm_kernel|att_q_kernel|att_qb_kernel' -s 105 -c 15 \
  -o gpurun_out/${TAG}_full -f $B > gpurun_out/${TAG}_full_ncu.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/${TAG}_full_ncu.log; ls -la gpurun_out/${TAG}_full.ncu-rep
