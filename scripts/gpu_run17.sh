set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum --launch-skip 1500 -c 2500 --csv --log-file gpurun_out/r02_launches_bench_decode.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-validate > gpurun_out/r02_launches_bench2.log 2>&1; tail -c 200 gpurun_out/r02_launches_bench2.log; wc -l gpurun_out/r02_launches_bench_decode.csv
