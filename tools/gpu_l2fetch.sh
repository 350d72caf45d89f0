#!/bin/bash
TAG=$1; shift
mkdir -p gpurun_out
for G in "$@"; do
  SCORE_L2_FETCH=$G timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_g$G.json 2> gpurun_out/${TAG}_g$G.err
  echo "gran $G rc=$?"; grep cudaLimit gpurun_out/${TAG}_g$G.err | head -1
  SCORE_L2_FETCH=$G timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'coatt_fwd_kernel|emb_update_kernel' -s 40 -c 4 --csv --log-file gpurun_out/${TAG}_g${G}_ncu.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > /dev/null 2>&1
  grep -v "^==" gpurun_out/${TAG}_g${G}_ncu.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Kernel Name'][:40], r['Metric Name'], r['Metric Value'], r['Metric Unit'])"
done
python tools/bench_digest.py gpurun_out/${TAG}_g*.json
