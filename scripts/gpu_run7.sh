set -x
mkdir -p gpurun_out
timeout 300 python scripts/dev_wa_trace.py 28672 4096 32 > gpurun_out/r2_trace_a.log 2>&1; cat gpurun_out/r2_trace_a.log | head -50
timeout 300 python scripts/dev_wa_trace.py 4096 4096 32 > gpurun_out/r2_trace_b.log 2>&1; cat gpurun_out/r2_trace_b.log | head -30
