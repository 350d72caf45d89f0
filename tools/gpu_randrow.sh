#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/bin/randrow_bench > gpurun_out/randrow_randrow.txt 2>&1; echo "randrow rc=$?"
grep -v "cpasync\|bulk\|depth [248]" gpurun_out/randrow_randrow.txt | tail -40
RANDROW_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none -k regex:"ldg_stride|rmw_kernel|ldg_kernel" --launch-skip 0 --launch-count 400 --csv --log-file gpurun_out/randrow_randrow_ncu.csv tools/bin/randrow_bench > /dev/null 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/randrow_randrow_ncu.csv
