set -x
mkdir -p gpurun_out
timeout 600 python scripts/dev_fa_tc.py > gpurun_out/r2_fa_tc.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2_fa_tc.log
timeout 600 python -m pytest tests/test_prefill_attn_gpu.py -q -x 2>&1 | tail -4
