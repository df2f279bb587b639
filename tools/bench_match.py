#!/usr/bin/env python
"""Measurement of match_contigs_b200 next to the reference's own loop nest.

One call = the whole matching (every single-copy gene one iteration) on the imputed matrix of a workload.  GPU: the drop-in
as a user calls it (imputed matrix uploaded once, one launch per iteration, networkx matching on the host).  CPU: the same
host logic with the edge weights from the reference's statements (oracle/binweights_oracle.bin_contact_weights_loops:
utility.py:466-480 as written, scalar LIL look-ups) - i.e. what the unmodified match_contigs does.  Prints one JSON line.
"""
import argparse, json, os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imputecc_b200 as pkg
from imputecc_b200 import synth
from oracle import binweights_oracle as bw
from oracle.match_inputs import marker_tables, smg_iterations

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2_sparse")
a = ap.parse_args()
wl = synth.WORKLOADS[a.workload]
M, idx = synth.make_workload(wl)
E = sp.csr_matrix(pkg.crwr_impute(M, idx, 0.5, 80, dtype=np.float32, device="cuda:0"))
E.sort_indices()
rev = {"contig%05d" % i: k for k, i in enumerate(idx)}
_, _, contig_markers = marker_tables(wl, idx)
smg = smg_iterations(contig_markers)
w_intra, w_inter = pkg.contact_thresholds(E, 50, 0)


class LoopsWeights:
    def __init__(self, E):
        self.E, self.seconds = E, 0.0

    def weights(self, to_bin, bin_ptr, members):
        t0 = time.perf_counter()
        W = bw.bin_contact_weights_loops(self.E, to_bin, bin_ptr, members)
        self.seconds += time.perf_counter() - t0
        return W


pkg.match_contigs_b200(smg, contig_markers, E, rev, w_intra, w_inter, device="cuda:0")      # warm-up (library, pinned buffers)
t0 = time.perf_counter()
got = pkg.match_contigs_b200(smg, contig_markers, E, rev, w_intra, w_inter, device="cuda:0")
gpu_s = time.perf_counter() - t0
ref_op = LoopsWeights(E)
t0 = time.perf_counter()
want = pkg.match_contigs_b200(smg, contig_markers, E, rev, w_intra, w_inter, _operator=ref_op)
cpu_s = time.perf_counter() - t0
assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
print(json.dumps({"metric": "match_contigs_calls_per_s", "unit": "calls/s", "workload": a.workload, "markers": int(idx.size),
                  "iterations": len(smg), "bins": int(got[2]), "value": 1.0 / gpu_s, "gpu_call_s": gpu_s,
                  "cpu_baseline": {"kind": "port", "cores": 1, "call_s": cpu_s, "edge_weight_loop_s": ref_op.seconds,
                                   "sample": "the whole call with the reference's statements for the edge weights (utility.py:466-480)"},
                  "speedup": cpu_s / gpu_s, "identical_bins": True}))
