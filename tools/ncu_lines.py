#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        hdr = r
    elif len(r) > 6 and hdr and r[0] not in ("", "Line No"):
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        num = lambda x: int(x) if x.isdigit() else 0
        lines.append((num(r[si]), num(r[ii]), cur_file, r[0], r[1].strip()))
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
for s, i, f, ln, src in sorted(lines, key=lambda l: -l[0])[:topn]:
    print("%6d %5.1f%% | inst %9d %5.1f%% | %s:%s | %s" % (s, 100.0 * s / tot_s, i, 100.0 * i / tot_i, f, ln, src[:100]))
