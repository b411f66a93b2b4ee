set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest3.log
tail -12 gpurun_out/r2_pytest3.log; grep -h "worst\|max err" gpurun_out/r2_pytest3.log
timeout 900 python scripts/dev_w4a16_perf.py > gpurun_out/r2_w4a16_perf3.log 2>&1
cat gpurun_out/r2_w4a16_perf3.log | tail -30
timeout 300 python scripts/dev_mmvq_perf.py > gpurun_out/r2_mmvq_perf3.log 2>&1
head -12 gpurun_out/r2_mmvq_perf3.log | cut -c1-140
timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 > gpurun_out/r2_sweep3.log; cat gpurun_out/r2_sweep3.log
