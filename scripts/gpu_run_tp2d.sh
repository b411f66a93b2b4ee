set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_gpu_check.py > gpurun_out/r2_tp2_check3.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_check3.log | tail -6
MRS_TP_LL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29621 scripts/dev_tp_breakdown.py > gpurun_out/r2_tp2_breakdown3.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_breakdown3.log | tail -4
MRS_TP_LL=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_tp2_ll.log 2> gpurun_out/r2_bench_tp2_ll.err; echo "rc=$?"; tail -c 600 gpurun_out/r2_bench_tp2_ll.log | head -c 400
