set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_model_gpu.py -q -s -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r2_pytest5.log; cat gpurun_out/r2_pytest5.log | tail -12
timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 | head -1 > gpurun_out/r2_sweep5a.log; cat gpurun_out/r2_sweep5a.log
MRS_DEV_LIB=$PWD/gpurun_old_lib.so timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 | head -1 > gpurun_out/r2_sweep5b.log; cat gpurun_out/r2_sweep5b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_int4 -s 3 -c 1 -f -o gpurun_out/r02_w4a16_int4 python scripts/dev_w4a16_one.py 28672 4096 32 > gpurun_out/r2_ncu_w4.log 2>&1; tail -3 gpurun_out/r2_ncu_w4.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; tail -c 7000 gpurun_out/r2_bench5.log; tail -5 gpurun_out/r2_bench5.err
