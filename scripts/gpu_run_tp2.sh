set -x
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_gpu_check.py > gpurun_out/r2_tp2_check.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_tp2_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_tp2.log 2> gpurun_out/r2_bench_tp2.err; echo "rc=$?"; tail -c 3000 gpurun_out/r2_bench_tp2.log; tail -8 gpurun_out/r2_bench_tp2.err
MRS_TP_NCCL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_tp2_nccl.log 2> gpurun_out/r2_bench_tp2_nccl.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_tp2_nccl.log; tail -5 gpurun_out/r2_bench_tp2_nccl.err
