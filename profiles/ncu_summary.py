#!/usr/bin/env python
"""Prints the handful of ncu metrics we track from a `--page raw --csv` export (one row per launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get('Kernel Name', '')[:100], "grid", d.get('Grid Size'), "block", d.get('Block Size'))
    for w in WANT:
        if w in d:
            print(f"  {w:82s} {d[w]:>16s} {units[hdr.index(w)]}")
    st = []
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(d[h].replace(',', '')), h.split('issue_stalled_')[1].split('_per_')[0]))
            except ValueError:
                pass
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
