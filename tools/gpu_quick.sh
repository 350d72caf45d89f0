#!/bin/bash
# quick single-GPU pass: parity tests + short bench digest.  Usage: tools/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-q}; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; else timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python tools/bench_digest.py gpurun_out/${TAG}_bench.json
