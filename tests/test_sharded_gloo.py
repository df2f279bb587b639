"""World-size-2 gloo test (CPU) of the sharded walk driver: run_sharded_walk + TorchComm move the
exchange buffers exactly as the GPU path does; the per-shard arithmetic is a numpy stand-in."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO
from oracle import crwr_oracle as oracle
from imputecc_b200 import sharded, synth
from imputecc_b200.crwr import shard_rows, stack_row_blocks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dtype_name, out_dir):
    import sys
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from fake_phases import NumpyPhases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dtype = np.dtype(dtype_name)
    M, idx = synth.make_workload(synth.Workload("g", 1200, 192, 60, 10))
    P = oracle.transition_matrix(M, oracle.marker_permutation(M.shape[0], idx), dtype)
    ranges = shard_rows(idx.size, world, weights=np.diff(M.indptr)[idx])
    ph = NumpyPhases(P, ranges[rank], 0.5, 80, 0.01, dtype)
    st = sharded.run_sharded_walk(ph, sharded.TorchComm(dist, dist.group.WORLD), 80, 500)
    sp.save_npz(os.path.join(out_dir, "q%d.npz" % rank), ph.Q.tocsr())
    np.save(os.path.join(out_dir, "t%d.npy" % rank), np.array([[t["nnz_in"], t["nnz_new"]] for t in ph.trace]))
    np.save(os.path.join(out_dir, "s%d.npy" % rank), np.array([st.steps_done, st.stop_reason]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_sharded_driver_world2_gloo(tmp_path, dtype_name):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, str(tmp_path)), nprocs=world, join=True)
    dtype = np.dtype(dtype_name)
    M, idx = synth.make_workload(synth.Workload("g", 1200, 192, 60, 10))
    P = oracle.transition_matrix(M, oracle.marker_permutation(M.shape[0], idx), dtype)
    tr = oracle.WalkTrace()
    want = oracle.crwr_walk(P, 0.5, 80, 0.01, idx.size, trace=tr, full_iterate=True)
    got = stack_row_blocks([sp.load_npz(str(tmp_path / ("q%d.npz" % r))) for r in range(world)])
    got.sort_indices(); want.sort_indices()
    s0, s1 = np.load(tmp_path / "s0.npy"), np.load(tmp_path / "s1.npy")
    assert np.array_equal(s0, s1) and s0[0] == len(tr.steps)          # same stop decision on both ranks
    t0 = np.load(tmp_path / "t0.npy")
    assert np.array_equal(t0, np.load(tmp_path / "t1.npy"))
    assert [tuple(x) for x in t0] == [(t.nnz_in, t.nnz_new) for t in tr.steps]
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=1e-6)
