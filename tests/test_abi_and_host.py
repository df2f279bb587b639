"""CPU tests: the C-ABI library loads and exports what include/crwr_b200.h declares; host logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import REPO
import imputecc_b200 as pkg
from imputecc_b200 import _lib, crwr, synth


def _declared_functions():
    text = open(os.path.join(REPO, "include", "crwr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crwr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build libcrwr_b200.so first (python -m imputecc_b200.build)"
    lib = C.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    # the ctypes binding covers exactly the header
    assert set(names) == set(_lib.exported_symbols())


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.crwr_abi_version() == 1
    assert isinstance(lib.crwr_last_error(), bytes)
    assert lib.crwr_launch_count(1) >= 0


def test_workspace_bytes_validates_config_without_gpu():
    lib = _lib.load()
    good = _lib.Config(_lib.CRWR_F32, 0, 1000, 100, 0, 100, 20000, 1 << 20, 1, 0)
    n = lib.crwr_workspace_bytes(C.byref(good))
    assert n > 2 * (1 << 20) * 8            # two iterate buffers
    bad = _lib.Config(7, 0, 1000, 100, 0, 100, 20000, 1 << 20, 1, 0)
    assert lib.crwr_workspace_bytes(C.byref(bad)) == -1
    assert b"dtype" in lib.crwr_last_error()
    bad_rows = _lib.Config(_lib.CRWR_F32, 0, 1000, 100, 0, 50, 20000, 1 << 20, 1, 0)   # unsharded must own all rows
    assert lib.crwr_workspace_bytes(C.byref(bad_rows)) == -1
    f64 = _lib.Config(_lib.CRWR_F64, 0, 1000, 100, 0, 100, 20000, 1 << 20, 1, 0)
    assert lib.crwr_workspace_bytes(C.byref(f64)) > n


def test_struct_layouts_match_header_sizes():
    assert C.sizeof(_lib.Config) == 64
    assert C.sizeof(_lib.StepRecord) == 56
    assert C.sizeof(_lib.Status) == 56


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    M = synth.contact_matrix(200, 20, 6, 1, seed=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.crwr_impute(M, np.arange(10), 0.5, 80)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.random_walk_b200(M.astype(np.float32), 0.5, 80, 0.01, 10)


def test_product_package_never_imports_the_oracle():
    for fn in os.listdir(os.path.join(REPO, "imputecc_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(REPO, "imputecc_b200", fn)).read()
            assert "oracle" not in src, fn


def test_markers_first_order_is_the_reference_permutation():
    idx = np.array([2, 5, 6])
    lst = list(idx)
    for i in range(9):                       # imputation.py:108-110
        if i not in lst:
            lst.append(i)
    new_of_old = crwr.markers_first_order(9, idx)
    assert np.array_equal(np.argsort(new_of_old), np.array(lst))
    assert new_of_old.dtype == np.int32


def test_validate_inputs_errors():
    M = synth.contact_matrix(100, 10, 4, 1, seed=1)
    ok = np.array([1, 4, 9])
    crwr.validate_inputs(M, ok, 0.5, 80)
    with pytest.raises(ValueError):
        crwr.validate_inputs(M[:50], ok, 0.5, 80)                    # not square
    with pytest.raises(ValueError):
        crwr.validate_inputs(M, np.array([], dtype=int), 0.5, 80)    # m = 0
    with pytest.raises(ValueError):
        crwr.validate_inputs(M, np.array([4, 1]), 0.5, 80)           # unsorted
    with pytest.raises(ValueError):
        crwr.validate_inputs(M, np.array([1, 100]), 0.5, 80)         # out of range
    with pytest.raises(ValueError):
        crwr.validate_inputs(M, ok, 0.5, 101)
    with pytest.raises(ValueError):
        crwr.validate_inputs(M, ok, 1.5, 80)
    neg = M.copy(); neg.data[0] = -1.0
    with pytest.raises(ValueError):
        crwr.validate_inputs(neg, ok, 0.5, 80)
    nan = M.copy(); nan.data[0] = np.nan
    with pytest.raises(ValueError):
        crwr.validate_inputs(nan, ok, 0.5, 80)


def test_shard_rows_partitions():
    for n, k in ((10, 1), (10, 3), (8000, 8), (9, 9)):
        parts = crwr.shard_rows(n, k)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(b > a for a, b in parts)
    w = np.r_[np.full(10, 100.0), np.full(90, 1.0)]
    parts = crwr.shard_rows(100, 2, weights=w)
    assert parts[0][1] <= 10                                         # balanced by weight, not by count
    with pytest.raises(ValueError):
        crwr.shard_rows(3, 4)


def test_expanding_walk_policy():
    """(steps issued, entries of the last product, cut) at the polls after walk steps 1 and 3, as recorded on the B200."""
    assert not crwr.expanding_walk((2, 533006, 8.104e-3), (4, 720825, 3.536e-3))          # c3_sparse, rwr-thres 80
    assert not crwr.expanding_walk((2, 10177701, 1.307e-3), (4, 19874412, 3.899e-4))      # c4_dense, rwr-thres 80
    assert crwr.expanding_walk((2, 533006, 1.509e-3), (4, 1626742, 1.405e-4))             # c3_sparse, rwr-thres 50
    assert not crwr.expanding_walk((2, 0, 0.0), (4, 10, 1e-3))                            # nothing to compare yet


def test_next_row_hints_policy():
    assert crwr.next_row_hints(0, 0, 80, 0) == (192, 32)
    out, kept = crwr.next_row_hints(30, 1, 80, 1)            # un-cut step-0 rows feed step 1
    assert kept == 30 and out >= 4 * 30
    out, kept = crwr.next_row_hints(1000, 30, 80, 2)         # first cut: kept share guessed from perc
    assert 1900 <= out <= 2100 and 500 <= kept <= 600
    assert crwr.next_row_hints(1000, 180, 80, 7) == (1000, 180)


def test_stack_row_blocks():
    A = sp.random(7, 5, density=0.4, format="csr", random_state=0)
    parts = [A[:3], A[3:4], A[4:]]
    assert (crwr.stack_row_blocks(parts) != A).nnz == 0


def test_synthetic_workloads_shape():
    M, idx = synth.make_workload(synth.Workload("t", 2000, 320, 100, 16))
    assert M.shape == (2000, 2000) and M.dtype == np.float64 and M.indices.dtype == np.int32
    assert (M != M.T).nnz == 0 and M.diagonal().sum() == 0 and M.has_sorted_indices
    assert idx.size == 320 and np.all(np.diff(idx) > 0)
    assert set(synth.WORKLOADS) >= {"c2_sparse", "c3_sparse", "c3_dense", "c4_sparse", "c4_dense"}


def test_install_rebinding_without_gpu():
    """install() only rebinds names; nothing runs until the reference calls them."""
    import types
    mod = types.ModuleType("stub")
    mod.random_walk_cpu = lambda *a: None
    mod.ImputeMatrix = type("ImputeMatrix", (), {"_imputation": lambda self: None})
    pkg.install(mod, "inner")
    assert mod.random_walk_cpu is pkg.random_walk_b200 and mod.ImputeMatrix._imputation is not pkg.imputation_b200
    pkg.install(mod, "outer")
    assert mod.ImputeMatrix._imputation is pkg.imputation_b200
    with pytest.raises(ValueError):
        pkg.install(mod, "both")
