"""Two ranks, ONE GPU: crwr_impute(group=WORLD) with two processes that both use cuda:0 and exchange the step
buffers through gloo (NCCL refuses two ranks on one device).  This is the whole public sharded call - row sharding,
the three collectives of a step, the regrow loop, the result gather, the symmetrisation - on a box with a single GPU,
so the round-end run covers it; the result must equal the unsharded one bit for bit."""
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


def _worker(rank, world, port, out_dir, iterate_cap):
    import sys
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import imputecc_b200 as pkg
    from imputecc_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M, idx = synth.make_workload(synth.Workload("n", 4000, 640, 100, 16))
    for name, dtype in (("f32", np.float32), ("f64", np.float64)):
        E, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=dtype, device=torch.device("cuda", 0), return_info=True,
                                  group=dist.group.WORLD, iterate_cap=iterate_cap)
        sp.save_npz(os.path.join(out_dir, "E_%s_%d.npz" % (name, rank)), E)
        np.save(os.path.join(out_dir, "info_%s_%d.npy" % (name, rank)), np.array([info.steps, info.regrows]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("iterate_cap,world", [(None, 2), (2000, 2), (None, 3)])
def test_ranks_on_one_gpu_match_single_walk(tmp_path, cuda_device, iterate_cap, world):
    import torch.multiprocessing as mp
    import imputecc_b200 as pkg
    from imputecc_b200 import synth
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), iterate_cap), nprocs=world, join=True)
    M, idx = synth.make_workload(synth.Workload("n", 4000, 640, 100, 16))
    for name, dtype in (("f32", np.float32), ("f64", np.float64)):
        want, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=dtype, device=cuda_device, return_info=True)
        want.sort_indices()
        for rank in range(world):
            got = sp.load_npz(str(tmp_path / ("E_%s_%d.npz" % (name, rank)))).tocsr()
            got.sort_indices()
            steps, regrows = np.load(tmp_path / ("info_%s_%d.npy" % (name, rank)))
            assert int(steps) == info.steps
            if iterate_cap is not None:
                assert int(regrows) >= 1                 # the undersized buffers were regrown on every rank, in lockstep
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert np.array_equal(got.data, want.data)
