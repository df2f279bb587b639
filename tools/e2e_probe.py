#!/usr/bin/env python
"""Developer probe: host-side time breakdown of crwr_impute()."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scipy.sparse as sp
import imputecc_b200 as pkg
from imputecc_b200 import crwr, synth, _lib
import ctypes as C

wl = sys.argv[1] if len(sys.argv) > 1 else "c3_sparse"
M, idx = synth.make_workload(synth.WORKLOADS[wl])
n, m = M.shape[0], idx.size
dev = torch.device("cuda:0")

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(4):
    t = [T()]
    crwr.validate_inputs(M, idx, 0.5, 80); t.append(T())
    Mc = sp.csr_matrix(M); _ = Mc.has_canonical_format; t.append(T())
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
print("workspace MB", w.cap, )
