#!/usr/bin/env python
"""Developer probe: cProfile of repeated crwr_impute() calls (host-side overhead of the end-to-end path)."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imputecc_b200 as pkg
from imputecc_b200 import synth
from bench import pinned_csr

wl = sys.argv[1] if len(sys.argv) > 1 else "c3_sparse"
M0, idx = synth.make_workload(synth.WORKLOADS[wl])
M = pinned_csr(M0)
dev = torch.device("cuda:0")
for _ in range(3):
    pkg.crwr_impute(M, idx, 0.5, 80, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pkg.crwr_impute(M, idx, 0.5, 80, device=dev)
torch.cuda.synchronize()
print("mean call %.3f ms" % ((time.perf_counter() - t0) * 100))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    pkg.crwr_impute(M, idx, 0.5, 80, device=dev)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
