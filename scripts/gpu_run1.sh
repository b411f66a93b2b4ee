set -x
mkdir -p gpurun_out
timeout 600 python tests/golden/make_ref_golden.py > gpurun_out/r2_golden.log 2>&1 && cp gpurun_out/ref_golden.npz tests/golden/ref_golden.npz
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest1.log
timeout 600 python scripts/dev_r2_sweep.py 32 > gpurun_out/r2_sweep1.log 2>&1
timeout 600 python scripts/dev_mmvq_perf.py > gpurun_out/r2_mmvq_perf1.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_sweep1.log; tail -3 gpurun_out/r2_bench1.log
