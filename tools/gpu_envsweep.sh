#!/bin/bash
# bench line per value of one environment knob.  Usage: tools/gpu_envsweep.sh <tag> <ENV_NAME> <values...>
TAG=$1; NAME=$2; shift; shift
mkdir -p gpurun_out
for C in "$@"; do
  env $NAME=$C timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_$C.json 2> gpurun_out/${TAG}_$C.err
  echo "$NAME=$C rc=$?"
done
python tools/bench_digest.py gpurun_out/${TAG}_*.json
