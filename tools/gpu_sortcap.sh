#!/bin/bash
# step time vs the CTA cap of the background sort.  Usage: tools/gpu_sortcap.sh <tag> <caps...>
TAG=$1; shift
mkdir -p gpurun_out
for C in "$@"; do
  SCORE_SORT_CTAS=$C timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_cap$C.json 2> gpurun_out/${TAG}_cap$C.err
  echo "cap $C rc=$?"
done
python tools/bench_digest.py gpurun_out/${TAG}_cap*.json
