set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_gpu_check.py > gpurun_out/r2_tp2_check2.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_check2.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29621 scripts/dev_tp_breakdown.py > gpurun_out/r2_tp2_breakdown2.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_breakdown2.log | tail -12
timeout 300 python scripts/dev_r2_sweep.py 32 2>&1 | head -1 > gpurun_out/r2_sweep12.log; cat gpurun_out/r2_sweep12.log
MRS_DEV_LIB=$PWD/gpurun_old_lib.so timeout 300 python scripts/dev_r2_sweep.py 32 2>&1 | head -1 > gpurun_out/r2_sweep12b.log; cat gpurun_out/r2_sweep12b.log
