#!/usr/bin/env python
"""Measurement of the bin-contact weights operator next to the reference's loop.

One "iteration" = the edge weights of one match_contigs iteration: |to_bin| contigs x |bins| bins over the imputed
matrix of a bench workload.  Prints one JSON line: GPU time through the public call (host buffers in, numpy out), kernel
time with CUDA events (inputs resident), achieved bytes/s against the measured HBM peak, and the reference's own
statements (oracle/binweights_oracle.bin_contact_weights_loops: scalar LIL look-ups) timed on a bounded sample.
"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imputecc_b200 as pkg
from imputecc_b200 import _lib, synth
from oracle import binweights_oracle as bw

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_sparse")
ap.add_argument("--to-bin", type=int, default=400)
ap.add_argument("--bins", type=int, default=500)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
M, idx = synth.make_workload(synth.WORKLOADS[a.workload])
E = sp.csr_matrix(pkg.crwr_impute(M, idx, 0.5, 80, dtype=np.float32, device="cuda:0"))
E.sort_indices()
m = E.shape[0]
rng = np.random.default_rng(3)
perm = rng.permutation(m)
to_bin = np.sort(perm[:a.to_bin]).astype(np.int32)
rest = perm[a.to_bin:]                                   # every other marker contig already sits in a bin
cuts = np.sort(rng.choice(np.arange(1, rest.size), size=a.bins - 1, replace=False))
bin_ptr = np.concatenate([[0], cuts, [rest.size]]).astype(np.int64)
members = rest.astype(np.int32)

for _ in range(3):
    W = pkg.bin_contact_weights(E, to_bin, bin_ptr, members, device="cuda:0")
t0 = time.perf_counter()
for _ in range(a.reps):
    W = pkg.bin_contact_weights(E, to_bin, bin_ptr, members, device="cuda:0")
e2e_ms = (time.perf_counter() - t0) / a.reps * 1e3
resident = pkg.BinContactWeights(E, device="cuda:0")           # as inside one match_contigs call: the matrix goes up once
for _ in range(3):
    W2 = resident.weights(to_bin, bin_ptr, members)
t0 = time.perf_counter()
for _ in range(a.reps):
    W2 = resident.weights(to_bin, bin_ptr, members)
e2e_resident_ms = (time.perf_counter() - t0) / a.reps * 1e3
assert np.array_equal(W2.view(np.uint32), W.view(np.uint32))

dev = torch.device("cuda:0")
up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
d = [up(E.indptr.astype(np.int64)), up(E.indices.astype(np.int32)), up(E.data), up(to_bin), up(bin_ptr), up(members)]
out = torch.empty((to_bin.size, a.bins), dtype=torch.float32, device=dev)
scratch = torch.empty(8 * m, dtype=torch.uint8, device=dev)
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def launch():
    _lib.check(lib.crwr_bin_contact_weights(0, m, C.c_void_p(d[0].data_ptr()), C.c_void_p(d[1].data_ptr()), C.c_void_p(d[2].data_ptr()),
                                            to_bin.size, C.c_void_p(d[3].data_ptr()), a.bins, C.c_void_p(d[4].data_ptr()),
                                            C.c_void_p(d[5].data_ptr()), members.size, C.c_void_p(scratch.data_ptr()), scratch.numel(),
                                            C.c_void_p(out.data_ptr()), st))
for _ in range(3): launch()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): launch()
e1.record(); torch.cuda.synchronize()
kern_us = e0.elapsed_time(e1) / a.reps * 1e3
assert np.array_equal(out.cpu().numpy().view(np.uint32), W.view(np.uint32))

# algorithmic bytes of one iteration: the rows of E of the contigs to bin, the member list and its inverse map once, the output
row_nnz = int(np.diff(E.indptr)[to_bin].sum())
alg_bytes = row_nnz * 8 + members.size * 4 + 2 * 8 * m + to_bin.size * a.bins * 4
peak = 6533.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
# CPU: the vectorised oracle on everything, the reference's statements on a bounded sample of the contigs
t0 = time.perf_counter(); Wo = bw.bin_contact_weights(E, to_bin, bin_ptr, members); vec_s = time.perf_counter() - t0
assert np.array_equal(Wo.view(np.uint32), W.view(np.uint32))
sample = to_bin[:4]
t0 = time.perf_counter(); bw.bin_contact_weights_loops(E, sample, bin_ptr, members); loops_s = (time.perf_counter() - t0) * to_bin.size / sample.size
print(json.dumps({"metric": "bin_contact_weights_iterations_per_s", "unit": "iterations/s", "workload": a.workload,
                  "contigs_to_bin": int(to_bin.size), "bins": a.bins, "binned_contigs": int(members.size), "m": m, "nnz_E": int(E.nnz),
                  "value": 1e6 / kern_us, "kernel_us": kern_us, "e2e_ms": e2e_ms, "e2e_matrix_resident_ms": e2e_resident_ms,
                  "roofline": {"bound": "hbm", "achieved": alg_bytes / (kern_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": alg_bytes / (kern_us * 1e-6) / 1e9 / peak, "bytes_per_launch": alg_bytes},
                  "cpu_baseline": {"kind": "port", "cores": 1, "reference_statements_s": loops_s,
                                   "sample": "utility.py:466-480 as written (scalar LIL look-ups) on %d of %d contigs, scaled" % (sample.size, to_bin.size),
                                   "vectorised_oracle_s": vec_s},
                  "speedup_e2e_vs_reference_statements": loops_s / (e2e_ms * 1e-3),
                  "speedup_matrix_resident_vs_reference_statements": loops_s / (e2e_resident_ms * 1e-3)}))
