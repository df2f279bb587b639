#!/usr/bin/env python
"""Developer probe: host-side time breakdown of crwr_impute() (same statements, timed)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scipy.sparse as sp
import imputecc_b200 as pkg
from imputecc_b200 import crwr, synth, _lib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import pinned_csr

wl = sys.argv[1] if len(sys.argv) > 1 else "c3_sparse"
M0, idx = synth.make_workload(synth.WORKLOADS[wl])
M = pinned_csr(M0)
n, m = M.shape[0], idx.size
dev = torch.device("cuda:0")

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(5):
    t = [T()]
    crwr.validate_inputs(M, idx, 0.5, 80, check_values=False); t.append(T())
    _ = M.has_canonical_format; t.append(T())
    noo = crwr.markers_first_order(n, idx); t.append(T())
    w = pkg.Walker(n, m, dtype=np.float32, device=dev, transition_nnz_cap=int(M.nnz) + n); t.append(T())
    w.build_transition(M.indptr, M.indices, M.data, noo); t.append(T())
    st = w.walk(0.5, 80); t.append(T())
    ip, ix, dt = w.result_device(); t.append(T())
    e_ip, e_ix, e_dt = crwr.symmetrise_device(ip, ix, dt, m); t.append(T())
    E = sp.csr_matrix((e_dt.cpu().numpy(), e_ix.cpu().numpy(), e_ip.cpu().numpy()), shape=(m, m)); t.append(T())
    w.close(); t.append(T())
    names = ["validate", "canonical", "order", "create", "transition", "walk", "result", "symmetrise", "download", "close"]
    print("rep %d total %.3f ms | " % (rep, 1e3 * (t[-1] - t[0])) + " ".join("%s %.3f" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t, t[1:])))
t0 = T(); E = pkg.crwr_impute(M, idx, 0.5, 80, device=dev); t1 = T()
print("crwr_impute call: %.3f ms" % (1e3 * (t1 - t0)))
