#!/usr/bin/env python
"""Aggregate executed instructions / stall samples of propagate_kernel by code phase (line ranges found from
marker comments in propagate.cuh)."""
import collections, csv, subprocess, sys, re, os
path = sys.argv[1]
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imputecc_b200/csrc/propagate.cuh")).read().splitlines()
def line_of(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise KeyError(pat)
k0 = line_of("__device__ __forceinline__ void propagate_rows(")
marks = [("setup", k0), ("stage_lambda", line_of("auto stage = [&]", k0)), ("row_fetch", line_of("WarpTotals wt;", k0)),
         ("row_load", line_of("// ---- read the row, dropping entries", k0)), ("sort+indptr", line_of("bool defer = nk > KMAX;", k0)),
         ("rounds_setup", line_of("// ---- the row's products in scipy's order", k0)), ("round_body", line_of("for (int r0 = 0; r0 < n_cur; r0 += 32) {", k0)),
         ("probe", line_of("if (leader) {", k0)), ("combine+finish_round", line_of("if (dup) {", k0)),
         ("after_rounds", line_of("if (defer) break;", k0)), ("epilogue", line_of("// ---- epilogue: scale, restart add, norm, write", k0)),
         ("end", line_of("totals_flush(a, wt);", k0) + 4)]
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None
agg = collections.Counter(); smp = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No": hdr = r
    elif len(r) > 6 and hdr and r[0] not in ("", "Line No"):
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        n = int(r[ii]) if r[ii].isdigit() else 0; s = int(r[si]) if r[si].isdigit() else 0
        ln = int(r[0])
        if cur == "propagate.cuh":
            ph = "other(propagate.cuh helpers)"
            for (name, a), (_, b) in zip(marks, marks[1:]):
                if a <= ln < b: ph = name
            if ln < k0:
                txt = src[ln - 1]
                ph = "helpers:" + ("sort" if 300 <= ln else "misc")
        else:
            ph = cur
        agg[ph] += n; smp[ph] += s
tn, ts = sum(agg.values()) or 1, sum(smp.values()) or 1
for k, v in agg.most_common():
    print("%-34s inst %10d %5.1f%%  samples %6d %5.1f%%" % (k, v, 100 * v / tn, smp[k], 100 * smp[k] / ts))
