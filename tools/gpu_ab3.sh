#!/bin/bash
# first GPU pass of the next round: parity + bench of the variants prepared without hardware access
#   SCORE_KE_SPLIT=1   tier S of the scatter/Adam kernel in a lean kernel of its own (scatter.cu)
# (the peer-memory exchange is tested by TRY_P2P=1 tools/gpu_multi.sh <tag> 2 under gpurun --gpus 2)
TAG=${1:-ab3}
mkdir -p gpurun_out
short() { python tools/bench_digest.py "$1" 2>/dev/null || cut -c1-300 "$1"; }
SCORE_KE_SPLIT=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_kesplit.log 2>&1; echo "pytest (KE_SPLIT) rc=$?"; tail -3 gpurun_out/${TAG}_pytest_kesplit.log
for V in 0 1 0 1; do
  SCORE_KE_SPLIT=$V timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_kesplit$V.json 2> gpurun_out/${TAG}_kesplit$V.err
  echo "KE_SPLIT=$V rc=$? $(short gpurun_out/${TAG}_kesplit$V.json)"
done
