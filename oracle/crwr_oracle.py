"""CPU oracle for ImputeCC's CRWR imputation path.  TEST INFRASTRUCTURE ONLY.

This file restates, with scipy/numpy on the CPU, the algorithm of
  * ``ImputeMatrix._imputation``   /root/reference/Script/imputation.py:93-129
  * ``random_walk_cpu``            /root/reference/Script/utility.py:270-310
so that the CUDA path in ``imputecc_b200`` has something to be checked against
on a box where ``/root/reference`` does not exist.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the product package never does.

Pinning.  The reference owns no golden vectors for this path (SURVEY.md §4,
§8c: "parity unpinned" by the reference's own tests).  The oracle is therefore
pinned against outputs of the *unmodified reference functions executed in the
build container*: ``oracle/make_golden.py`` imports ``Script.utility`` /
``Script.imputation`` from ``/root/reference``, runs them on the bundled
``test_data`` matrix and on small synthetic matrices, and freezes inputs and
outputs under ``tests/golden/`` (numpy 2.3.5 / scipy 1.18.1; the reference pins
numpy 1.21.6 / scipy 1.7.3 in ImputeCC_env.yaml:9,12, which are not installable
here).  ``tests/test_oracle.py`` checks this restatement against those files
bit for bit.

The arithmetic itself lives in third-party code that is not under
``/root/reference`` (scipy.sparse ``_sparsetools``: csr_matmat, csr_plus_csr,
csr_minus_csr, csr_elmul_csr; numpy: partition + linear interpolation).  The
oracle calls the same library entry points in the same order, because the
order fixes the floating-point rounding; a numpy-only restatement of the
percentile (``percentile_linear``) and of one propagation step
(``propagate_dense``) is provided to cross-check the device kernels
independently of scipy.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import norm as sparse_norm

# hard-wired constants of the reference loop
MAX_STEPS = 500            # utility.py:281
PLATEAU_EPS = 0.001        # utility.py:298
PLATEAU_COUNT = 5          # utility.py:302
DEFAULT_TOL = 0.01         # imputation.py:120


@dataclass
class StepTrace:
    """What one pass of utility.py:282-304 did."""
    step: int
    nnz_in: int        # stored entries of the thresholded iterate entering the step
    nnz_new: int       # stored entries of (1-rp) Q P + rp I
    nnz_kept: int      # stored entries after the percentile cut
    cut: float         # percentile value (nan when the cut is skipped: step 0)
    cut_applied: bool
    delta: float       # ||Q_prev - Q_new||_F
    plateau_hits: int  # cumulative count of |delta - delta_prev| < 0.001


@dataclass
class WalkTrace:
    steps: List[StepTrace] = field(default_factory=list)
    stop_reason: str = "max_steps"      # "tolerance" | "plateau" | "max_steps"
    seconds: float = 0.0


def marker_permutation(n_contigs: int, marker_idx_sorted) -> np.ndarray:
    """Markers first (ascending), every other contig after (ascending).

    imputation.py:101 builds the sorted marker list and :108-110 appends the
    rest with an O(N^2) ``not in`` scan; the resulting list is this vector.
    """
    marker_idx_sorted = np.asarray(marker_idx_sorted, dtype=np.int64)
    rest = np.ones(n_contigs, dtype=bool)
    rest[marker_idx_sorted] = False
    return np.concatenate([marker_idx_sorted, np.nonzero(rest)[0]])


def transition_matrix(normcc: sp.spmatrix, perm: np.ndarray, dtype=np.float32) -> sp.csr_matrix:
    """Permuted, diagonal-free, self-looped, row-scaled walk matrix P.

    imputation.py:112-119.  Row i of B is multiplied by 1/colsum_i(B); the
    result is cast to ``dtype`` (the reference casts to float32; float64 gives
    the fp64 parity oracle of SURVEY.md §8c without touching the algorithm).
    """
    perm = list(perm)
    shuffled = normcc.copy().tocsr()[perm, :]
    shuffled = shuffled.tocsc()[:, perm]                              # :112-113
    A = shuffled.copy()
    A = A - sp.diags(A.diagonal())                                    # :116
    isolated = (np.asarray(A.sum(axis=0)).ravel() == 0).astype(int)
    B = A + sp.diags(isolated)                                        # :117
    scale = sp.diags(1 / np.asarray(B.sum(axis=0)).ravel())           # :118
    return scale.dot(B).astype(dtype)                                 # :119


def crwr_walk(P, rp, perc, tol, num_marker_contig, max_steps: int = MAX_STEPS,
              trace: Optional[WalkTrace] = None, full_iterate: bool = False):
    """Constrained random walk with restart, utility.py:270-310.

    Returns the m x m CSC slice (utility.py:306-307); with ``full_iterate`` the
    m x N CSR iterate after the last cut instead.  When ``trace`` is given it is
    filled with one StepTrace per step; tracing does not change any value.
    """
    m = num_marker_contig
    t0 = time.perf_counter()
    ones = np.ones(m)
    ar = np.arange(m)
    restart = sp.coo_matrix((ones, (ar, ar)), shape=(m, P.shape[0]), dtype=np.float32)  # :278
    Q = restart.copy()
    delta_prev = 0
    plateau = 0
    reason = "max_steps"
    for it in range(max_steps):
        nnz_in = Q.nnz
        Q_new = (1 - rp) * Q.dot(P) + rp * restart                    # :282
        delta = sparse_norm(Q - Q_new)                                # :283
        Q = Q_new.copy()                                              # :284
        nnz_new = Q.nnz
        cut = np.nan
        applied = False
        if it >= 1:                                                   # :286
            cut = np.percentile(Q.tocoo().data, perc)                 # :287
            if cut > 0:                                               # :288
                Q = Q.multiply(Q > cut)                               # :290
                applied = True
        stop = False
        if delta < tol:                                               # :293
            reason = "tolerance"
            stop = True
        else:
            if np.abs(delta - delta_prev) < PLATEAU_EPS:              # :298
                plateau += 1
            delta_prev = delta                                        # :301
            if plateau == PLATEAU_COUNT:                              # :302
                reason = "plateau"
                stop = True
        if trace is not None:
            trace.steps.append(StepTrace(it, int(nnz_in), int(nnz_new), int(Q.nnz), float(cut),
                                         applied, float(delta), plateau))
        if stop:
            break
    if trace is not None:
        trace.stop_reason = reason
        trace.seconds = time.perf_counter() - t0
    if full_iterate:
        return Q.tocsr()
    Q = Q.tocsr()[np.arange(m), :]                                    # :306
    Q = Q.tocsc()[:, np.arange(m)]                                    # :307
    return Q


def symmetrise_normalise(Q) -> sp.spmatrix:
    """E = Q + Q^T, then D^-1/2 E D^-1/2 with d = colsum(E), zeros -> 1.

    imputation.py:122-127 (the reference then calls ``.tolil()``, :129).
    """
    E = Q.copy()
    E += E.T                                                          # :123
    d = np.asarray(E.sum(axis=0)).ravel()                             # :124
    d[d == 0] = 1                                                     # :125
    b = sp.diags(1 / np.sqrt(d))                                      # :126
    return b.dot(E).dot(b)                                            # :127


def crwr_impute(normcc: sp.spmatrix, marker_idx_sorted, rp: float, perc, tol: float = DEFAULT_TOL,
                dtype=np.float32, trace: Optional[WalkTrace] = None, max_steps: int = MAX_STEPS):
    """B-outer of SURVEY.md §8b: NormCC matrix + marker index set -> imputed m x m matrix."""
    marker_idx_sorted = np.asarray(marker_idx_sorted)
    perm = marker_permutation(normcc.shape[0], marker_idx_sorted)
    P = transition_matrix(normcc, perm, dtype)
    Q = crwr_walk(P, rp, perc, tol, len(marker_idx_sorted), max_steps=max_steps, trace=trace)
    return symmetrise_normalise(Q)


# ---------------------------------------------------------------------------
# scipy-free restatements used to cross-check single device kernels
# ---------------------------------------------------------------------------

def percentile_linear(values: np.ndarray, perc) -> float:
    """numpy's method='linear' percentile, restated on a full sort.

    Follows numpy/lib/_function_base_impl.py (2.3.5): virtual index (n-1)q,
    ``_get_indexes``, ``_get_gamma`` and ``_lerp``.  numpy evaluates the
    virtual index in the dtype of the data (float32 data -> float32 index), so
    this does too; see SURVEY.md §8c "oracle version hazard".
    """
    x = np.sort(np.asarray(values))
    n = x.shape[0]
    ft = x.dtype.type
    q = np.true_divide(perc, ft(100))            # percentile(): divisor takes the data dtype
    v = (n - 1) * q                              # _QuantileMethods['linear']['get_virtual_index']
    k = np.floor(v)
    gamma = v - k                                # _get_gamma, 'linear' leaves it unchanged
    if v >= n - 1:                               # _get_indexes: clamp to the last element
        lo = hi = n - 1
    elif v < 0:
        lo = hi = 0
    else:
        lo, hi = int(k), int(k) + 1
    a, b = x[lo], x[hi]
    diff = b - a                                 # _lerp
    out = a + diff * gamma
    if gamma >= 0.5:
        out = b - diff * (1 - gamma)
    return ft(out)


def propagate_dense(Qd: np.ndarray, P: sp.csr_matrix, rp: float) -> np.ndarray:
    """(1-rp) Q P + rp I on a dense m x N iterate (small cases only)."""
    m = Qd.shape[0]
    out = (1 - rp) * (Qd @ P.toarray().astype(Qd.dtype))
    out[np.arange(m), np.arange(m)] += rp
    return out


def step_bytes(nnz_p: int, n: int, nnz_in: int, nnz_new: int, nnz_kept: int, s: int,
               first_step: bool = False) -> int:
    """Algorithmic bytes of one walk step, SURVEY.md §8d.

    B_step = nP(s+4) + 4(N+1) + (s+4)(nI + nO + nK) + s nO; step 0 omits the
    selection read and the kept write.
    """
    b = nnz_p * (s + 4) + 4 * (n + 1) + (s + 4) * (nnz_in + nnz_new)
    if not first_step:
        b += (s + 4) * nnz_kept + s * nnz_new
    return b
