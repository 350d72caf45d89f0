#!/bin/bash
# hop2 GPU tests, compute-sanitizer (memcheck + racecheck) on the smoke script, Zipf 1.05 bench line
mkdir -p gpurun_out
T=r2i
timeout 600 python -m pytest tests/test_hop2.py -m gpu -q > gpurun_out/${T}_hop2.log 2>&1; echo "hop2 rc=$?"; tail -5 gpurun_out/${T}_hop2.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${T}_racecheck.log
timeout 300 python bench.py --steps 200 --warmup 10 --zipf 1.05 --no-cpu-baseline > gpurun_out/${T}_zipf105.json 2> gpurun_out/${T}_zipf105.err; echo "zipf rc=$?"
python tools/bench_digest.py gpurun_out/${T}_zipf105.json
