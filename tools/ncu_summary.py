#!/usr/bin/env python
"""Per-launch summary of an ncu report (the counters DESIGN.md / profiles/ quote).
usage: ncu_summary.py <report.ncu-rep> [tokens_total]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
tokens = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second']
tot_inst = tot_ms = tot_dram = 0.0
for n, r in enumerate(rows[2:], 1):
    print(f"== launch {n} of {len(rows) - 2} ==")
    for w in want:
        if w in idx:
            print(f"{w} [{units[idx[w]]}] = {r[idx[w]]}")
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            v = float(r[idx[h]])
            if v > 0.25:
                print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {v:.3f}")
    tot_inst += float(r[idx['smsp__inst_executed.sum']])
    tot_ms += float(r[idx['gpu__time_duration.sum']]) * (1e-3 if units[idx['gpu__time_duration.sum']] == 'us' else 1.0)
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        u = units[idx[k]]
        tot_dram += float(r[idx[k]]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
    print()
print(f"TOTAL: {tot_ms:.3f} ms (serialised), {tot_inst:.4g} warp instructions, {tot_dram / 1e9:.3f} GB DRAM")
if tokens:
    print(f"per token: {tot_inst / tokens:.1f} warp instructions, {tot_dram / tokens:.0f} DRAM bytes")
