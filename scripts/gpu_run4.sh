set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest4.log
tail -8 gpurun_out/r2_pytest4.log; grep -h "worst\|max err" gpurun_out/r2_pytest4.log
timeout 900 python scripts/dev_w4a16_perf.py > gpurun_out/r2_w4a16_perf4.log 2>&1
grep -v "M=  1\|M=  8" gpurun_out/r2_w4a16_perf4.log | tail -20
timeout 600 python scripts/dev_r2_sweep.py 32 2>&1 > gpurun_out/r2_sweep4.log; cat gpurun_out/r2_sweep4.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err; tail -c 6000 gpurun_out/r2_bench4.log; tail -5 gpurun_out/r2_bench4.err
