set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29621 scripts/dev_tp_breakdown.py > gpurun_out/r2_tp2_breakdown.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_tp2_breakdown.log | tail -12
