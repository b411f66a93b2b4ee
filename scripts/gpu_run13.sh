set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mmq_gpu.py -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r2_pytest13.log; tail -8 gpurun_out/r2_pytest13.log
timeout 900 python scripts/dev_mmq_perf.py > gpurun_out/r2_mmq_perf13.log 2>&1; cat gpurun_out/r2_mmq_perf13.log | tail -20
