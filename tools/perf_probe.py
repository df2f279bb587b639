#!/usr/bin/env python
"""Developer probe: one workload, a few walks, per-step trace and kernel timing (not the bench)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imputecc_b200 as pkg
from imputecc_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_sparse")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--hints", default="")
ap.add_argument("--rp", type=float, default=0.5)
ap.add_argument("--perc", type=float, default=80)
ap.add_argument("--custom", default="", help="n_contigs,n_markers,genome_size,intra_degree instead of a named workload")
args = ap.parse_args()
dt = np.float64 if args.dtype == "f64" else np.float32
if args.custom:
    M, idx = synth.make_workload(synth.Workload("custom", *[int(x) for x in args.custom.split(",")]))
else:
    M, idx = synth.make_workload(synth.WORKLOADS[args.workload])
n, m = M.shape[0], idx.size
w = pkg.Walker(n, m, dtype=dt, device="cuda:0", transition_nnz_cap=M.nnz + n)
w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(n, idx))
if args.hints:
    w.hint_override = tuple(int(x) for x in args.hints.split(","))
lib = _lib.load()
for rep in range(args.reps):
    lib.crwr_profile(w.handle, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = w.walk(args.rp, args.perc)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tot, nl = _lib.C.c_double(0), _lib.C.c_int64(0)
    lib.crwr_profile_read(w.handle, _lib.C.byref(tot), _lib.C.byref(nl), w._stream())
    ph = (_lib.C.c_double * 3)()
    lib.crwr_profile_phases(w.handle, ph)
    print("rep %d: %d steps, wall %.3f ms, propagate total %.3f ms over %d launches | phases: prop+dense %.3f hist1 %.3f gather+finish %.3f" % (rep, st.steps_done, wall * 1e3, tot.value, nl.value, ph[0], ph[1], ph[2]))
shape = (_lib.C.c_int64 * 8)()
lib.crwr_get_shape(w.handle, shape)
print("shape H %d limit %d kmax %d scap %d warps %d per_warp %d grid %d sms %d" % tuple(shape))
gs = (_lib.C.c_int64 * 8)()
lib.crwr_get_group_shape(w.handle, gs)
print("group shape H %d limit %d kmax %d ov %d G %d warps %d per_row %d grid %d" % tuple(gs))
for t in w.trace():
    print("step %2d in %8d new %9d maxrow %5d maxkept %5d fallback %5d fast %6d cut %.3e delta %.4f" %
          (t["step"], t["nnz_in"], t["nnz_new"], t["max_row_len"], t["max_kept"], t["fallback_rows"], t["fast_rows"], t["cut"], t["delta"]))
print("retries", w.retries, "regrows", w.regrows)
w.close()
