set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r2_pytest14.log; tail -6 gpurun_out/r2_pytest14.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench14.log 2> gpurun_out/r2_bench14.err; tail -c 9000 gpurun_out/r2_bench14.log; tail -5 gpurun_out/r2_bench14.err
