"""Two processes, two GPUs, real NCCL: crwr_impute(group=WORLD) must return the unsharded result bit for bit.
Skipped unless the box has at least two CUDA devices (the round-end `pytest -m gpu` run uses one)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, env):
    import sys
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import imputecc_b200 as pkg
    from imputecc_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.update(env)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    M, idx = synth.make_workload(synth.Workload("n", 6000, 960, 100, 16))
    for name, dtype in (("f32", np.float32), ("f64", np.float64)):
        E, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=dtype, device=torch.device("cuda", rank), return_info=True,
                                  group=dist.group.WORLD)
        sp.save_npz(os.path.join(out_dir, "E_%s_%d.npz" % (name, rank)), E)
        np.save(os.path.join(out_dir, "steps_%s_%d.npy" % (name, rank)), np.array([info.steps]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("env", [{},                                                       # peer path: window steps (one sync per step)
                                 {"CRWR_WIN_MIN": "1e-7", "CRWR_WIN_SCALE": "0.01"},      # windows that miss: steps repeated in lockstep
                                 {"CRWR_WIN_STEP": "0"},                                   # three flag barriers per step
                                 {"CRWR_SHARDED_EXCHANGE": "nccl"}])                       # NCCL collectives
def test_two_gpu_nccl_matches_single_gpu(tmp_path, env):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import imputecc_b200 as pkg
    from imputecc_b200 import synth
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), env), nprocs=2, join=True)
    M, idx = synth.make_workload(synth.Workload("n", 6000, 960, 100, 16))
    for name, dtype in (("f32", np.float32), ("f64", np.float64)):
        want, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=dtype, device="cuda:0", return_info=True)
        want.sort_indices()
        for rank in range(2):
            got = sp.load_npz(str(tmp_path / ("E_%s_%d.npz" % (name, rank)))).tocsr()
            got.sort_indices()
            assert int(np.load(tmp_path / ("steps_%s_%d.npy" % (name, rank)))[0]) == info.steps
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert np.array_equal(got.data, want.data)
