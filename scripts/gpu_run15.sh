set -x
mkdir -p gpurun_out
timeout 600 python scripts/dev_fa_tc.py > gpurun_out/r2_fa_tc.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_fa_tc.log
