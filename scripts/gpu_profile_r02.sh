# round-2 profile pass (one GPU): the launch list of the bench command + one `ncu --set full` capture per hot kernel
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-validate > gpurun_out/r02_launches_bench.log 2>&1; tail -c 300 gpurun_out/r02_launches_bench.log
timeout 600 $NCU --set full -k regex:mmvq_stream --launch-skip 42 -c 5 -f -o gpurun_out/r02_mmvq_decode python scripts/profile_targets.py decode > gpurun_out/r02_p1.log 2>&1; tail -2 gpurun_out/r02_p1.log
timeout 600 $NCU --set full -k regex:paged_decode --launch-skip 8 -c 2 -f -o gpurun_out/r02_attn_decode python scripts/profile_targets.py decode > gpurun_out/r02_p2.log 2>&1; tail -2 gpurun_out/r02_p2.log
timeout 600 $NCU --set full -k regex:mmq_ts --launch-skip 2 -c 1 -f -o gpurun_out/r02_mmq_ts_q8_0 python scripts/profile_targets.py mmq q8_0 > gpurun_out/r02_p3.log 2>&1; tail -2 gpurun_out/r02_p3.log
timeout 600 $NCU --set full -k regex:mmq_ts --launch-skip 2 -c 1 -f -o gpurun_out/r02_mmq_ts_q4_k python scripts/profile_targets.py mmq q4_k > gpurun_out/r02_p3b.log 2>&1; tail -2 gpurun_out/r02_p3b.log
timeout 600 $NCU --set full -k regex:w4a16_int4 --launch-skip 3 -c 1 -f -o gpurun_out/r02_w4a16_int4 python scripts/profile_targets.py w4a16 > gpurun_out/r02_p4.log 2>&1; tail -2 gpurun_out/r02_p4.log
timeout 600 $NCU --set full -k regex:prefill_attn_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r02_prefill_attn python scripts/profile_targets.py prefill_attn mma > gpurun_out/r02_p5.log 2>&1; tail -2 gpurun_out/r02_p5.log
timeout 600 $NCU --set full -k regex:prefill_attn_tc --launch-skip 1 -c 1 -f -o gpurun_out/r02_prefill_attn_tc python scripts/profile_targets.py prefill_attn > gpurun_out/r02_p6.log 2>&1; tail -2 gpurun_out/r02_p6.log
for f in gpurun_out/r02_*.ncu-rep; do python scripts/ncu_summary.py $f > ${f%.ncu-rep}.summary.txt 2>&1; done
# (reports that embed the whole module are too big to bring back: keep their summaries)
for f in gpurun_out/r02_attn_decode gpurun_out/r02_mmvq_decode; do ncu -i $f.ncu-rep --page details --csv > $f.details.csv 2>/dev/null; rm -f $f.ncu-rep; done
ls -la gpurun_out/
du -sh gpurun_out
