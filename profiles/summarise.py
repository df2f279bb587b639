#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_X.csv  > profiles/rN_X_launches.txt
    python profiles/summarise.py kernel   gpurun_out/prof_X.ncu-rep  > profiles/rN_X_kernel.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void crwr::", "").replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none : %d launches, %.1f us total" % (len(rows), tot / 1e3))
    print("# per-launch times are cold-cache and serialised: compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s n=%4d total=%10.1f us mean=%9.2f us share=%5.1f%%" % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / tot))


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-82s %s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
