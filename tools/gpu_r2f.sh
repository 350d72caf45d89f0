#!/bin/bash
mkdir -p gpurun_out
for V in 0 1; do
SCORE_COATT_LEAN=$V timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "full_size_taobao_values" > gpurun_out/r2f_full_lean$V.log 2>&1; echo "lean=$V rc=$?"
grep -c "OUTSIDE" gpurun_out/r2f_full_lean$V.log; grep "above" gpurun_out/r2f_full_lean$V.log | head -8
done
