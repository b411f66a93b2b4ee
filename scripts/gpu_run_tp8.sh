set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_tp8.log 2> gpurun_out/r2_bench_tp8.err; echo "rc=$?"; tail -c 2500 gpurun_out/r2_bench_tp8.log; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_tp8.err | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29632 bench.py --config 5 --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_cfg5.log 2> gpurun_out/r2_bench_cfg5.err; echo "rc=$?"; tail -c 2500 gpurun_out/r2_bench_cfg5.log; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_cfg5.err | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r2_bench_tp4.log 2> gpurun_out/r2_bench_tp4.err; echo "rc=$?"; tail -c 800 gpurun_out/r2_bench_tp4.log
