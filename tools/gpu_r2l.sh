#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 120 --csv --log-file gpurun_out/r2l_shard1_launches.csv python tools/shard_world1.py large_vocab_shard 6 > gpurun_out/r2l_shard1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2l_shard1.log
python tools/launch_summary.py gpurun_out/r2l_shard1_launches.csv 2>/dev/null | head -50
