"""Sharded walk on ONE GPU: K shards driven in lockstep in one process, the collectives replaced by
in-process sums/concatenation of the same exchange buffers (B200_PROFILING.md: emulate ranks as one
program when there are fewer GPUs than ranks).  Exercises the crwr_step_* entry points."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import load_csr, load_golden, same_support
from oracle import crwr_oracle as oracle
import imputecc_b200 as pkg
from imputecc_b200 import sharded, synth

pytestmark = pytest.mark.gpu


class MultiPhases:
    def __init__(self, phases):
        self.p = phases

    def exchange(self, which):
        return [p.exchange(which) for p in self.p]

    def __getattr__(self, name):
        if name in ("propagate", "hist", "scan", "gather", "set_row_hints"):
            return lambda *a: [getattr(p, name)(*a) for p in self.p]
        raise AttributeError(name)

    def finish(self, gathered):
        for p in self.p:
            p.finish(gathered)

    def poll(self):
        sts = [p.poll() for p in self.p]
        assert len({(s.steps_done, s.stop_reason) for s in sts}) == 1      # identical decisions on every shard
        return sts[0]

    def row_hints(self, st, perc, issued):
        hs = [p.row_hints(p.poll(), perc, issued) for p in self.p]
        return max(h[0] for h in hs), max(h[1] for h in hs)


class LoopbackComm:
    def all_reduce_sum(self, views):
        total = torch.stack(views).sum(0)
        for v in views:
            v.copy_(total)

    def all_gather(self, blocks):
        return torch.cat(blocks)

    def max_ints(self, values, like):
        return tuple(values)


@pytest.mark.parametrize("shards", [2, 3])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sharded_walk_equals_unsharded(cuda_device, shards, dtype):
    M, idx = synth.make_workload(synth.Workload("s", 3000, 480, 100, 16))
    n, m = M.shape[0], idx.size
    new_of_old = pkg.markers_first_order(n, idx)
    single = pkg.Walker(n, m, dtype=dtype, device=cuda_device, transition_nnz_cap=M.nnz + n)
    single.build_transition(M.indptr, M.indices, M.data, new_of_old)
    st1 = single.walk(0.5, 80)
    want = single.iterate()
    trace1 = single.trace()
    single.close()

    ranges = pkg.shard_rows(m, shards, weights=np.diff(M.indptr)[idx])
    walkers = [pkg.Walker(n, m, dtype=dtype, device=cuda_device, transition_nnz_cap=M.nnz + n, row_range=r, n_shards=shards)
               for r in ranges]
    try:
        for w in walkers:
            w.build_transition(M.indptr, M.indices, M.data, new_of_old)
            w.walk_init(0.5, 80)
        st = sharded.run_sharded_walk(MultiPhases([sharded.CudaPhases(w) for w in walkers]), LoopbackComm(), 80, 500)
        got = pkg.stack_row_blocks([w.iterate() for w in walkers])
        traces = [w.trace() for w in walkers]
    finally:
        for w in walkers:
            w.close()
    assert st.steps_done == st1.steps_done and st.stop_reason == st1.stop_reason
    for tr in traces:                                              # every shard logged the global numbers
        assert [(t["nnz_in"], t["nnz_new"], t["cut"], t["delta"]) for t in tr[1:]] == \
               [(t["nnz_in"], t["nnz_new"], t["cut"], t["delta"]) for t in trace1[1:]]
    got.sort_indices(); want.sort_indices()
    assert same_support(got, want) and np.array_equal(got.data, want.data)   # bit-identical, any sharding

