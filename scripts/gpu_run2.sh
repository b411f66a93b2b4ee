set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_w4a16_gpu.py tests/test_gptq_model_gpu.py tests/test_ref_golden_gpu.py tests/test_gptq_gpu.py -q -x -s 2>&1 | tail -40 > gpurun_out/r2_pytest2a.log
tail -15 gpurun_out/r2_pytest2a.log
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest2.log
tail -8 gpurun_out/r2_pytest2.log; grep -h "worst\|max err" gpurun_out/r2_pytest2.log gpurun_out/r2_pytest2a.log
timeout 900 python scripts/dev_w4a16_perf.py > gpurun_out/r2_w4a16_perf.log 2>&1
cat gpurun_out/r2_w4a16_perf.log | tail -30
timeout 300 python scripts/dev_mmvq_perf.py > gpurun_out/r2_mmvq_perf2.log 2>&1
timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 | head -2 > gpurun_out/r2_sweep2.log; cat gpurun_out/r2_sweep2.log
