set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -8 > gpurun_out/r2_pytest16.log; tail -5 gpurun_out/r2_pytest16.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench16.log 2> gpurun_out/r2_bench16.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench16.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['frac'], d['prefill'], d['config4']['hnd'], d['cpu_baseline']['value'], d['gpu_reference']['decode_tok_s'])
PY
tail -3 gpurun_out/r2_bench16.err
timeout 600 ncu --clock-control none --set full -k regex:prefill_attn_tc --launch-skip 1 -c 1 -f -o gpurun_out/r02_prefill_attn_tc python scripts/profile_targets.py prefill_attn > gpurun_out/r02_p6.log 2>&1; python scripts/ncu_summary.py gpurun_out/r02_prefill_attn_tc.ncu-rep > gpurun_out/r02_prefill_attn_tc.summary.txt; head -30 gpurun_out/r02_prefill_attn_tc.summary.txt; ls -la gpurun_out/r02_prefill_attn_tc.ncu-rep
