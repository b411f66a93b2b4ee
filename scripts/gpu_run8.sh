set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_w4a16_gpu.py tests/test_gptq_model_gpu.py tests/test_ref_golden_gpu.py tests/test_ops_gpu.py -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r2_pytest12.log; tail -6 gpurun_out/r2_pytest12.log
timeout 900 python scripts/dev_w4a16_perf.py > gpurun_out/r2_w4a16_perf12.log 2>&1; tail -30 gpurun_out/r2_w4a16_perf12.log
timeout 300 python scripts/dev_wa_trace.py 28672 4096 32 > gpurun_out/r2_trace12a.log 2>&1; cat gpurun_out/r2_trace12a.log | head -40
timeout 300 python scripts/dev_wa_trace.py 4096 4096 32 > gpurun_out/r2_trace12b.log 2>&1; cat gpurun_out/r2_trace12b.log | head -30
