set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_w4a16_gpu.py tests/test_gptq_model_gpu.py tests/test_sampler_gpu.py tests/test_mmvq_gpu.py tests/test_ref_golden_gpu.py -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r2_pytest6.log; tail -12 gpurun_out/r2_pytest6.log
timeout 900 python scripts/dev_w4a16_perf.py > gpurun_out/r2_w4a16_perf6.log 2>&1; tail -30 gpurun_out/r2_w4a16_perf6.log
timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 | head -1 > gpurun_out/r2_sweep6.log; cat gpurun_out/r2_sweep6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_int4 -s 3 -c 1 -f -o gpurun_out/r02_w4a16_int4_b python scripts/dev_w4a16_one.py 28672 4096 32 > gpurun_out/r2_ncu_w4b.log 2>&1; tail -3 gpurun_out/r2_ncu_w4b.log
timeout 600 python scripts/gpu_reference_chain.py > gpurun_out/r2_refchain6.log 2>&1; tail -8 gpurun_out/r2_refchain6.log
