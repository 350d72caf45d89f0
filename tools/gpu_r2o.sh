#!/bin/bash
mkdir -p gpurun_out
T=r2o
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "wide_rows or forward_backward_small or ragged" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for TH in 256 512 1024; do
SCORE_GRU_THREADS=$TH timeout 300 python bench.py --workload large_vocab_shard --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${T}_lv_$TH.json 2> gpurun_out/${T}_lv_$TH.err; echo "lv threads=$TH rc=$?"
python tools/bench_digest.py gpurun_out/${T}_lv_$TH.json
done
for TH in 256 512 1024; do
SCORE_GRU_THREADS=$TH timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 20 > gpurun_out/${T}_tb_$TH.json 2> gpurun_out/${T}_tb_$TH.err; echo "taobao threads=$TH rc=$?"
python tools/bench_digest.py gpurun_out/${T}_tb_$TH.json
done
SCORE_GRU_THREADS=1024 timeout 300 python bench.py --workload ccmr --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${T}_ccmr_1024.json 2> gpurun_out/${T}_ccmr_1024.err; python tools/bench_digest.py gpurun_out/${T}_ccmr_1024.json
