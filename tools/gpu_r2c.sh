#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/bin/randrow_bench > gpurun_out/r2c_randrow.txt 2>&1; echo "randrow rc=$?"
cat gpurun_out/r2c_randrow.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest.log
