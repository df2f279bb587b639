import glob
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_csr(rec, prefix):
    shape = tuple(int(x) for x in rec[prefix + "_shape"])
    return sp.csr_matrix((rec[prefix + "_data"], rec[prefix + "_indices"], rec[prefix + "_indptr"]), shape=shape)


def golden_names(pattern="*"):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, pattern + ".npz"))
                  if "constructor" not in p)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def same_support(A, B):
    A = sp.csr_matrix(A); B = sp.csr_matrix(B)
    A.sort_indices(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
