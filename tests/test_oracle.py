"""The CPU oracle against outputs of the unmodified reference (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_names, load_csr, load_golden, same_support
from oracle import crwr_oracle as oracle


def _bit_equal(A, B):
    A = sp.csr_matrix(A); B = sp.csr_matrix(B)
    A.sort_indices(); B.sort_indices()
    return same_support(A, B) and A.dtype == B.dtype and np.array_equal(A.data, B.data)


@pytest.mark.parametrize("name", golden_names())
def test_transition_matrix_bit_identical(name):
    rec = load_golden(name)
    normcc = load_csr(rec, "normcc")
    perm = oracle.marker_permutation(normcc.shape[0], rec["marker_idx"])
    P = oracle.transition_matrix(normcc, perm, np.float32)
    assert _bit_equal(P, load_csr(rec, "P32"))


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("tag,dtype", [("32", np.float32), ("64", np.float64)])
def test_walk_and_imputed_matrix_bit_identical(name, tag, dtype):
    rec = load_golden(name)
    normcc = load_csr(rec, "normcc")
    m = len(rec["marker_idx"])
    perm = oracle.marker_permutation(normcc.shape[0], rec["marker_idx"])
    # golden walk64 = unmodified random_walk_cpu fed the reference's float32 P upcast to float64
    P = oracle.transition_matrix(normcc, perm, np.float32).astype(dtype)
    tr = oracle.WalkTrace()
    Q = oracle.crwr_walk(P, float(rec["rp"]), int(rec["perc"]), 0.01, m, trace=tr)
    assert Q.dtype == dtype and Q.format == "csc"
    assert _bit_equal(Q, load_csr(rec, "walk" + tag))
    E = oracle.symmetrise_normalise(Q)
    assert _bit_equal(E, load_csr(rec, "E" + tag))
    # trace bookkeeping is self-consistent
    assert tr.steps[0].nnz_in == m and not tr.steps[0].cut_applied
    for a, b in zip(tr.steps, tr.steps[1:]):
        assert b.nnz_in == a.nnz_kept
    assert tr.stop_reason in ("plateau", "tolerance", "max_steps")


def test_constructor_path_matches_injected_path():
    """ImputeMatrix driven through its real constructor == the injected-input run (same markers)."""
    a = load_golden("bundled_constructor_m120_t80_rp50")
    b = load_golden("bundled_m120_t80_rp50")
    assert np.array_equal(a["marker_idx"], b["marker_idx"])
    assert _bit_equal(load_csr(a, "E32"), load_csr(b, "E32"))


def test_expected_nnz_table():
    """final nnz of Q[:m,:m], fp32, rp 0.5."""
    want = {(60, 50): 724, (120, 50): 1883, (250, 50): 4846, (60, 80): 158, (120, 80): 282, (250, 80): 384,
            (60, 95): 0, (120, 95): 0, (250, 95): 0}
    for (m, t), nnz in want.items():
        rec = load_golden("bundled_m%d_t%d_rp50" % (m, t))
        assert load_csr(rec, "walk32").nnz == nnz


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 99_991])
@pytest.mark.parametrize("perc", [0, 37, 50, 80, 95, 100])
def test_percentile_restatement_matches_numpy(dtype, n, perc):
    x = np.random.default_rng(n + perc).random(n).astype(dtype)
    x[:: 7] = x[0]                                    # ties
    got = oracle.percentile_linear(x, perc)
    want = np.percentile(x, perc)
    assert got.dtype == want.dtype and got == want


def test_marker_permutation_matches_reference_list_semantics():
    idx = np.array([3, 7, 8])
    lst = list(idx)
    for i in range(12):                               # imputation.py:108-110
        if i not in lst:
            lst.append(i)
    assert np.array_equal(oracle.marker_permutation(12, idx), np.array(lst))


def test_step_bytes_formula():
    # c3 sparse steady state, fp32 (16.7 MB)
    b = oracle.step_bytes(431_650, 50_000, 171_000, 856_000, 171_000, 4)
    assert abs(b / 1e6 - 16.7) < 0.5
