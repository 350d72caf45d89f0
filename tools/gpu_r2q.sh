#!/bin/bash
mkdir -p gpurun_out
T=r2q
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --zipf 1.05 --no-cpu-baseline > gpurun_out/${T}_zipf105.json 2> gpurun_out/${T}_zipf105.err; echo "zipf rc=$?"
python tools/bench_digest.py gpurun_out/${T}_zipf105.json
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_taobao.json 2> gpurun_out/${T}_taobao.err; echo "taobao rc=$?"
python tools/bench_digest.py gpurun_out/${T}_taobao.json
