"""Summarise an ncu report (read on the CPU box): python scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for k in want:
        for i, h in enumerate(hdr):
            if h == k or h.endswith('.' + k):
                print(f"  {k:72s} {r[i]:>14s} {units[i]}")
                break
    st = []
    for i, h in enumerate(hdr):
        if 'average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
            try: st.append((float(r[i]), h.split('issue_stalled_')[1].replace('.ratio','')))
            except Exception: pass
    st.sort(reverse=True)
    print('  stalls (warp latency cycles per issue):', ', '.join(f"{n}={v:.2f}" for v, n in st[:8]))
