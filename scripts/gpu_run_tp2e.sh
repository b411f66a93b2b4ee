set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_gpu_check.py > gpurun_out/r2_tp2_check4.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_check4.log | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_tp2_final.log 2> gpurun_out/r2_bench_tp2_final.err; echo "rc=$?"; tail -c 500 gpurun_out/r2_bench_tp2_final.log | head -c 300; grep -i "bench\]" gpurun_out/r2_bench_tp2_final.err | tail -3
