#!/usr/bin/env python
"""Headline metrics + per-source-line hot spots of one .ncu-rep (needs -lineinfo, --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, vals = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for h, v in zip(hdr, vals):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.3):
        print("%-90s %s" % (h, v))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, L = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No": hdr = r
    elif len(r) > 6 and hdr and r[0].isdigit():
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        num = lambda x: int(x) if x.isdigit() else 0
        L.append((num(r[si]), num(r[ii]), cur, int(r[0]), r[1].strip()))
ts = sum(l[0] for l in L) or 1; ti = sum(l[1] for l in L) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
for s, i, f, ln, src in sorted(L, key=lambda l: -l[1])[:topn]:
    print("%5.1f%% smp | %9d %5.1f%% inst | %s:%d | %s" % (100.0 * s / ts, i, 100.0 * i / ti, f, ln, src[:95]))
