"""GPU parity tests: the CUDA path, called through the public API / C ABI, against the oracle and
the golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py).

Gates (north_star): float64 walk - identical support, values within rel 1e-6 of the reference's
float64 run; float32 walk (the reference's own arithmetic) - same gate against the reference's
float32 run.  Because the kernels keep scipy's summation order, both are in fact expected to be
BIT-identical; `test_bit_identity_*` pins that stronger property separately.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import golden_names, load_csr, load_golden, same_support
from oracle import crwr_oracle as oracle
import imputecc_b200 as pkg
from imputecc_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6        # north_star: "matches the reference within rel 1e-6"


def _canon(A):
    A = sp.csr_matrix(A).copy()
    A.sort_indices()
    return A


def _assert_close(got, want, what):
    got, want = _canon(got), _canon(want)
    assert got.shape == want.shape, what
    assert same_support(got, want), "%s: support differs (nnz %d vs %d)" % (what, got.nnz, want.nnz)
    if want.nnz:
        rel = np.abs(got.data.astype(np.float64) - want.data.astype(np.float64)) / np.abs(want.data.astype(np.float64))
        assert rel.max() <= REL_TOL, "%s: max rel err %.3e" % (what, rel.max())


def _bit_equal(got, want):
    got, want = _canon(got), _canon(want)
    return same_support(got, want) and got.dtype == want.dtype and np.array_equal(got.data, want.data)


def _P_from_golden(rec, dtype):
    normcc = load_csr(rec, "normcc")
    perm = oracle.marker_permutation(normcc.shape[0], rec["marker_idx"])
    return oracle.transition_matrix(normcc, perm, np.float32).astype(dtype)     # as make_golden fed it


# ---------------------------------------------------------------------------------------------
# K1: transition matrix
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["bundled_m120_t80_rp50", "synth1500_t80_rp50", "synth3000dense_t80_rp50", "asym400_t60_rp50"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_transition_matrix(cuda_device, name, dtype):
    rec = load_golden(name)
    normcc = load_csr(rec, "normcc")
    idx = rec["marker_idx"]
    w = pkg.Walker(normcc.shape[0], len(idx), dtype=dtype, device=cuda_device, transition_nnz_cap=normcc.nnz + normcc.shape[0])
    try:
        w.build_transition(normcc.indptr, normcc.indices, normcc.data, pkg.markers_first_order(normcc.shape[0], idx))
        P = w.transition()
    finally:
        w.close()
    want = oracle.transition_matrix(normcc, oracle.marker_permutation(normcc.shape[0], idx), dtype)
    assert P.dtype == dtype and P.has_sorted_indices
    assert _bit_equal(P, want)
    if dtype == np.float32:
        assert _bit_equal(P, load_csr(rec, "P32"))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_transition_matrix_many_tiles(cuda_device, dtype):
    """N large enough for the multi-CTA offset scans of the build (tiles of 8192 contigs)."""
    M, idx = synth.make_workload(synth.Workload("t20k", 20000, 3200, 100, 16))
    w = pkg.Walker(M.shape[0], len(idx), dtype=dtype, device=cuda_device, transition_nnz_cap=M.nnz + M.shape[0])
    try:
        w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(M.shape[0], idx))
        P = w.transition()
    finally:
        w.close()
    want = oracle.transition_matrix(M, oracle.marker_permutation(M.shape[0], idx), dtype)
    assert P.dtype == dtype and P.has_sorted_indices
    assert _bit_equal(P, want)


# ---------------------------------------------------------------------------------------------
# B-inner: random_walk_b200 == random_walk_cpu
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("tag,dtype", [("64", np.float64), ("32", np.float32)])
def test_walk_matches_reference(cuda_device, name, tag, dtype):
    rec = load_golden(name)
    P = _P_from_golden(rec, dtype)
    m = len(rec["marker_idx"])
    Q, info = pkg.random_walk_b200(P, float(rec["rp"]), int(rec["perc"]), 0.01, m, device=cuda_device, return_info=True)
    want = load_csr(rec, "walk" + tag)
    assert Q.format == "csc" and Q.dtype == dtype and Q.shape == (m, m)
    # per-step trace against the oracle first: it localises a failure to a step
    tr = oracle.WalkTrace()
    oracle.crwr_walk(P, float(rec["rp"]), int(rec["perc"]), 0.01, m, trace=tr)
    got_steps = [(t["nnz_in"], t["nnz_new"]) for t in info.trace]
    want_steps = [(t.nnz_in, t.nnz_new) for t in tr.steps]
    assert got_steps == want_steps, "trace diverges: %r vs %r" % (got_steps[:6], want_steps[:6])
    assert info.stop_reason == tr.stop_reason
    for g, t in zip(info.trace, tr.steps):
        assert abs(g["delta"] - t.delta) <= 1e-5 * max(1.0, t.delta)
        if t.step >= 1:
            assert g["cut"] == pytest.approx(t.cut, rel=1e-6)
    _assert_close(Q, want, name + "/walk" + tag)


@pytest.mark.parametrize("name", ["bundled_m120_t80_rp50", "bundled_m250_t50_rp50", "synth1500_t80_rp50", "synth1500_t50_rp30",
                                  "synth3000dense_t80_rp50", "asym400_t60_rp50", "bundled_m120_t80_rp30"])
@pytest.mark.parametrize("tag,dtype", [("64", np.float64), ("32", np.float32)])
@pytest.mark.parametrize("structures", ["auto", "always"])
def test_bit_identity_with_reference_walk(cuda_device, name, tag, dtype, structures, monkeypatch):
    """Both propagation forms - the hash-accumulating row-group kernel and the cached-structure kernels (symbolic /
    numeric split, which rows this short only take when forced) - against the vectors of the unmodified reference."""
    if structures == "always":
        monkeypatch.setenv("CRWR_STRUCT", "2")
    rec = load_golden(name)
    P = _P_from_golden(rec, dtype)
    Q = pkg.random_walk_b200(P, float(rec["rp"]), int(rec["perc"]), 0.01, len(rec["marker_idx"]), device=cuda_device)
    assert _bit_equal(Q, load_csr(rec, "walk" + tag))


# ---------------------------------------------------------------------------------------------
# B-outer: crwr_impute == ImputeMatrix._imputation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_impute_matches_reference_float32(cuda_device, name):
    rec = load_golden(name)
    normcc = load_csr(rec, "normcc")
    E = pkg.crwr_impute(normcc, rec["marker_idx"], float(rec["rp"]), int(rec["perc"]), dtype=np.float32, device=cuda_device)
    assert E.dtype == np.float32
    _assert_close(E, load_csr(rec, "E32"), name + "/E32")


@pytest.mark.parametrize("name", ["bundled_m120_t80_rp50", "synth1500_t80_rp50", "synth3000dense_t80_rp50", "asym400_t60_rp50"])
def test_impute_float64_matches_oracle(cuda_device, name):
    rec = load_golden(name)
    normcc = load_csr(rec, "normcc")
    E = pkg.crwr_impute(normcc, rec["marker_idx"], float(rec["rp"]), int(rec["perc"]), dtype=np.float64, device=cuda_device)
    want = oracle.crwr_impute(normcc, rec["marker_idx"], float(rec["rp"]), int(rec["perc"]), dtype=np.float64)
    _assert_close(E, want, name + "/E64")


def test_symmetrise_normalise_bit_identical(cuda_device):
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        Q = sp.random(300, 300, density=0.05, random_state=rng, format="csr").astype(dtype)
        Q.data = (Q.data + dtype(0.01)).astype(dtype)
        Q[5, :] = 0; Q[:, 5] = 0; Q.eliminate_zeros()            # an empty row/column: d == 0 -> 1
        got = pkg.symmetrise_normalise(Q, device=cuda_device)
        want = oracle.symmetrise_normalise(Q.tocsc())
        assert _bit_equal(got, want)


def test_drop_in_for_ImputeMatrix(cuda_device):
    """imputation_b200 reads the same attributes and returns the same tuple as _imputation."""
    rec = load_golden("bundled_m120_t80_rp50")
    normcc = load_csr(rec, "normcc")
    n = normcc.shape[0]

    class Obj:
        pass
    o = Obj()
    names = np.array(["ctg%06d" % i for i in range(n)], dtype=object)
    o.contig_info = np.stack([names, np.zeros(n, dtype=object), np.ones(n, dtype=object)], axis=1)
    o.dict_contigRev = {nm: i for i, nm in enumerate(names)}
    o.normcc_matrix = normcc
    o.contig_markers = {names[i]: ["g"] for i in rec["marker_idx"][::-1]}       # any dict order
    o.rwr_rp, o.rwr_perc = 0.5, 80
    local, rev, E = pkg.imputation_b200(o, device=cuda_device)
    assert sp.isspmatrix_lil(E) and E.dtype == np.float32
    assert list(local[:, 0]) == list(names[rec["marker_idx"]])
    assert rev[names[rec["marker_idx"][3]]] == 3
    _assert_close(E.tocsr(), load_csr(rec, "E32"), "drop-in E")
    import pickle
    assert pickle.loads(pickle.dumps(E)).nnz == E.nnz                           # ImputeCC.py:241 pickles the result


# ---------------------------------------------------------------------------------------------
# larger live-oracle cases and size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c2_sized_walk_against_live_oracle(cuda_device, dtype):
    M, idx = synth.make_workload(synth.Workload("c2ish", 6000, 960, 100, 16))
    E, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=dtype, device=cuda_device, return_info=True)
    tr = oracle.WalkTrace()
    want = oracle.crwr_impute(M, idx, 0.5, 80, dtype=dtype, trace=tr)
    assert info.steps == len(tr.steps) and info.stop_reason == tr.stop_reason
    assert [t["nnz_new"] for t in info.trace] == [t.nnz_new for t in tr.steps]
    _assert_close(E, want, "c2ish")
    Es = sp.csr_matrix(E)
    asym = abs(Es - Es.T)                       # (b_i e) b_j vs (b_j e) b_i: symmetric up to rounding, as in the reference
    assert (Es != Es.T).nnz == 0 or asym.max() <= 4 * np.finfo(dtype).eps * Es.max()


@pytest.mark.parametrize("perc", [50, 80, 95])
@pytest.mark.parametrize("rp", [0.3, 0.5, 0.7])
def test_parameter_sweep_against_live_oracle(cuda_device, rp, perc):
    """BASELINE config 5 (rwr-rp x rwr-thres sweep) at a size the oracle walks in a fraction of a second: the imputed
    matrix is bit-identical in float32, so anything computed from it downstream (preliminary bins) is identical too."""
    M, idx = synth.make_workload(synth.Workload("sweep", 4000, 640, 100, 16))
    E, info = pkg.crwr_impute(M, idx, rp, perc, dtype=np.float32, device=cuda_device, return_info=True)
    tr = oracle.WalkTrace()
    want = oracle.crwr_impute(M, idx, rp, perc, dtype=np.float32, trace=tr)
    assert info.steps == len(tr.steps) and info.stop_reason == tr.stop_reason
    assert [(t["nnz_in"], t["nnz_new"]) for t in info.trace] == [(t.nnz_in, t.nnz_new) for t in tr.steps]
    assert _bit_equal(sp.csr_matrix(E), sp.csr_matrix(want))


@pytest.mark.parametrize("name", ["c2_sparse", "c3_sparse", "c3_dense", "c4_sparse", "c4_dense"])
def test_baseline_sized_workloads_bit_identical_to_live_oracle(cuda_device, name):
    """The bench workloads themselves (BASELINE configs 2, 3 and 4 - 32,000 markers on 200,000 contigs - SURVEY.md 8d
    generator), float32 as the reference computes: every kernel shape the bench runs against the scipy walk on the
    same matrix.  The oracle takes ~0.5 / 2 / 11 / 10 / 45 s."""
    M, idx = synth.make_workload(synth.WORKLOADS[name])
    E, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=np.float32, device=cuda_device, return_info=True)
    tr = oracle.WalkTrace()
    want = oracle.crwr_impute(M, idx, 0.5, 80, dtype=np.float32, trace=tr)
    assert info.steps == len(tr.steps) and info.stop_reason == tr.stop_reason
    assert [(t["nnz_in"], t["nnz_new"]) for t in info.trace] == [(t.nnz_in, t.nnz_new) for t in tr.steps]
    # the sizing hints keep the settled walk off the large-row path (c4_dense: a handful of its longest rows may take it)
    assert sum(t["fallback_rows"] for t in info.trace[4:]) <= (8 if name == "c4_dense" else 0)
    assert _bit_equal(sp.csr_matrix(E), sp.csr_matrix(want))


def test_c3_sparse_float64_gate(cuda_device):
    """north_star's gate at the headline size: float64 walk of the 8,000-marker workload against the float64 run of the
    same algorithm - identical support, values within rel 1e-6 (they are in fact bit-identical)."""
    M, idx = synth.make_workload(synth.WORKLOADS["c3_sparse"])
    E, info = pkg.crwr_impute(M, idx, 0.5, 80, dtype=np.float64, device=cuda_device, return_info=True)
    tr = oracle.WalkTrace()
    want = oracle.crwr_impute(M, idx, 0.5, 80, dtype=np.float64, trace=tr)
    assert info.steps == len(tr.steps) and info.stop_reason == tr.stop_reason
    assert [(t["nnz_in"], t["nnz_new"]) for t in info.trace] == [(t.nnz_in, t.nnz_new) for t in tr.steps]
    _assert_close(E, want, "c3_sparse/E64")
    assert _bit_equal(sp.csr_matrix(E), sp.csr_matrix(want))


@pytest.mark.parametrize("perc", [50, 80, 95])
@pytest.mark.parametrize("rp", [0.3, 0.5, 0.7])
def test_parameter_sweep_at_8000_markers(cuda_device, rp, perc):
    """BASELINE config 5 as stated: rwr-rp {0.3,0.5,0.7} x rwr-thres {50,80,95} at 8,000 markers (c3_sparse).  The
    imputed matrix is bit-identical to the reference algorithm's, so everything computed from it downstream - the
    preliminary bins of PreCluster included - is identical as well.  At thres 50 the scipy walk keeps half of a
    fast-growing iterate and needs minutes per cell, so those three cells compare the first 7 walk steps (both sides
    stop at max_steps; every kernel path has run by then)."""
    M, idx = synth.make_workload(synth.WORKLOADS["c3_sparse"])
    max_steps = 7 if perc == 50 else 500
    E, info = pkg.crwr_impute(M, idx, rp, perc, max_steps=max_steps, dtype=np.float32, device=cuda_device, return_info=True)
    tr = oracle.WalkTrace()
    want = oracle.crwr_impute(M, idx, rp, perc, dtype=np.float32, trace=tr, max_steps=max_steps)
    assert info.steps == len(tr.steps) and (info.stop_reason == tr.stop_reason or max_steps == 7)
    assert [(t["nnz_in"], t["nnz_new"]) for t in info.trace] == [(t.nnz_in, t.nnz_new) for t in tr.steps]
    assert _bit_equal(sp.csr_matrix(E), sp.csr_matrix(want))


def test_result_is_independent_of_kernel_sizing_and_large_row_path(cuda_device):
    """Any accumulator sizing - including one so small that most rows take the dense large-row
    kernel - must give the same bits."""
    rec = load_golden("synth3000dense_t80_rp50")
    P = _P_from_golden(rec, np.float64)
    m = len(rec["marker_idx"])
    outs = []
    for hints in (None, (16, 4), (300, 40), (2000, 2000), (40000, 30000)):
        w = pkg.Walker(P.shape[0], m, dtype=np.float64, device=cuda_device, transition_nnz_cap=P.nnz)
        try:
            w.set_transition(P)
            w.hint_override = hints
            st = w.walk(0.5, 80)
            outs.append((int(st.steps_done), int(st.fallback_rows), w.iterate()))
        finally:
            w.close()
    assert outs[1][1] > 0                         # the tiny sizing really exercised the large-row path
    for steps, _, it in outs[1:]:
        assert steps == outs[0][0]
        assert _bit_equal(it, outs[0][2])


def test_iterate_rows_sum_below_one_and_restart_present(cuda_device):
    """Properties that hold at any size: rows of the kept iterate are sub-stochastic for a symmetric
    input, and every value exceeds the last cut."""
    M, idx = synth.make_workload(synth.Workload("p", 4000, 640, 100, 16))
    n, m = M.shape[0], idx.size
    w = pkg.Walker(n, m, dtype=np.float64, device=cuda_device, transition_nnz_cap=M.nnz + n)
    try:
        w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(n, idx))
        P = w.transition()
        assert np.allclose(np.asarray(P.sum(axis=1)).ravel(), 1.0, atol=1e-12)   # symmetric input: row-stochastic
        st = w.walk(0.5, 80)
        it = w.iterate()
        assert it.shape == (m, n)
        assert np.all(np.asarray(it.sum(axis=1)).ravel() <= 1.0 + 1e-9)
        assert it.data.min() > st.last_cut
        tr = w.trace()
        assert tr[-1]["nnz_new"] >= it.nnz and len(tr) == st.steps_done
    finally:
        w.close()


def test_empty_result_and_edge_percentiles(cuda_device):
    rec = load_golden("bundled_m120_t95_rp50")
    E = pkg.crwr_impute(load_csr(rec, "normcc"), rec["marker_idx"], 0.5, 95, device=cuda_device)
    assert E.nnz == 0 and E.shape == (120, 120)                                 # pre_clustering.py:47 branches on this
    M = synth.contact_matrix(300, 30, 8, 1, seed=9)
    one = pkg.crwr_impute(M, np.array([17]), 0.5, 80, device=cuda_device)      # m = 1
    assert one.shape == (1, 1)
    want = oracle.crwr_impute(M, np.array([17]), 0.5, 80)
    _assert_close(one, want, "m=1")


def test_argument_errors(cuda_device):
    M = synth.contact_matrix(300, 30, 8, 1, seed=9)
    with pytest.raises(ValueError):
        pkg.crwr_impute(M, np.array([], dtype=int), 0.5, 80, device=cuda_device)
    with pytest.raises(ValueError):
        pkg.crwr_impute(M, np.array([3, 2]), 0.5, 80, device=cuda_device)
    with pytest.raises(ValueError):
        pkg.crwr_impute(M, np.array([3]), 0.5, 180, device=cuda_device)
    with pytest.raises(ValueError):
        pkg.random_walk_b200(M.astype(np.float32), 0.5, 80, 0.01, 0, device=cuda_device)


def test_capacity_regrow(cuda_device):
    """An undersized iterate buffer is detected on the device and the walk restarts larger."""
    rec = load_golden("synth1500_t50_rp30")
    P = _P_from_golden(rec, np.float32)
    m = len(rec["marker_idx"])
    w = pkg.Walker(P.shape[0], m, dtype=np.float32, device=cuda_device, transition_nnz_cap=P.nnz, iterate_cap=4 * m)
    try:
        w.set_transition(P)
        w.walk(0.3, 50)
        assert w.regrows >= 1
        assert _bit_equal(w.result().tocsc(), load_csr(rec, "walk32"))
    finally:
        w.close()


def test_bad_contact_values_are_rejected_on_the_device(cuda_device):
    M = synth.contact_matrix(300, 30, 8, 1, seed=9)
    for bad in (-1.0, np.nan, np.inf):
        B = M.copy()
        B.data[7] = bad
        with pytest.raises(ValueError, match="finite and non-negative"):
            pkg.crwr_impute(B, np.array([3, 9, 40]), 0.5, 80, device=cuda_device)


def _stub_imputation_module():
    """A stand-in for the imported `Script.imputation`: the statements of imputation.py that matter to install()."""
    import types
    mod = types.ModuleType("stub_imputation")

    def random_walk_cpu(P, rp, perc, tol, m):
        raise AssertionError("the CPU walk must have been rebound")

    class ImputeMatrix:
        def __init__(self, normcc, contig_info, contig_markers, rp, perc):
            self.normcc_matrix, self.contig_info, self.contig_markers = normcc, contig_info, contig_markers
            self.dict_contigRev = {nm: i for i, nm in enumerate(contig_info[:, 0])}
            self.rwr_rp, self.rwr_perc = rp, perc
            self.contig_local, self.dict_contigRevLocal, self.imputed_matrix = self._imputation()      # imputation.py:89

        def _imputation(self):
            # imputation.py:98-129 with the oracle's pre/post steps around the module-level walk (as the reference does)
            idx = np.sort(np.array([self.dict_contigRev[c] for c in self.contig_markers], dtype=np.int64))
            perm = oracle.marker_permutation(self.normcc_matrix.shape[0], idx)
            P = oracle.transition_matrix(self.normcc_matrix, perm, np.float32)
            Q = mod.random_walk_cpu(P, self.rwr_rp, self.rwr_perc, 0.01, len(idx))
            local = self.contig_info[idx, :]
            return local, {nm: i for i, nm in enumerate(local[:, 0])}, oracle.symmetrise_normalise(Q).tolil()

    mod.random_walk_cpu = random_walk_cpu
    mod.ImputeMatrix = ImputeMatrix
    return mod


@pytest.mark.parametrize("mode", ["outer", "inner"])
def test_install_rebinds_the_reference_entry_points(cuda_device, mode):
    """install(Script.imputation): "outer" replaces ImputeMatrix._imputation, "inner" the name random_walk_cpu that
    imputation.py:9 imported; constructing ImputeMatrix then runs the GPU path and yields the reference's matrix."""
    rec = load_golden("bundled_m120_t80_rp50")
    normcc = load_csr(rec, "normcc")
    n = normcc.shape[0]
    names = np.array(["ctg%06d" % i for i in range(n)], dtype=object)
    info = np.stack([names, np.zeros(n, dtype=object), np.ones(n, dtype=object)], axis=1)
    markers = {names[i]: ["g"] for i in rec["marker_idx"]}
    mod = _stub_imputation_module()
    pkg.install(mod, mode)
    if mode == "outer":
        assert mod.ImputeMatrix._imputation is pkg.imputation_b200
    else:
        assert mod.random_walk_cpu is pkg.random_walk_b200
    imp = mod.ImputeMatrix(normcc, info, markers, 0.5, 80)
    assert sp.isspmatrix_lil(imp.imputed_matrix) and imp.imputed_matrix.dtype == np.float32
    assert list(imp.contig_local[:, 0]) == list(names[rec["marker_idx"]])
    assert _bit_equal(imp.imputed_matrix.tocsr(), load_csr(rec, "E32"))
    with pytest.raises(ValueError):
        pkg.install(mod, "sideways")
    if mode == "outer":
        # the CSR hand-off (imputecc_b200/handoff.py): the same matrix without the LIL conversion, and every call the
        # clustering stages make on it still works
        mod2 = _stub_imputation_module()
        pkg.install(mod2, "outer", hand_off="csr")
        imp2 = mod2.ImputeMatrix(normcc, info, markers, 0.5, 80)
        assert sp.isspmatrix_csr(imp2.imputed_matrix) and _bit_equal(imp2.imputed_matrix, load_csr(rec, "E32"))
        assert np.array_equal(np.sort(imp2.imputed_matrix.tocoo().data), np.sort(imp.imputed_matrix.tocoo().data))
        assert _bit_equal(imp2.imputed_matrix.tolil().tocsr(), imp.imputed_matrix.tocsr())
        assert _bit_equal(imp2.imputed_matrix.copy().tocsr()[[0, 3, 5], :], imp.imputed_matrix.copy().tocsr()[[0, 3, 5], :])
        assert pkg.storage_offenders(imp2) == [] and pkg.storage_offenders(imp) == []


# ---------------------------------------------------------------------------------------------
# cached row structures (structure.cuh) and window select (tail.cuh)
# ---------------------------------------------------------------------------------------------
def _walk(M, idx, dtype, device, walker_cls=None):
    n, m = M.shape[0], idx.size
    w = (walker_cls or pkg.Walker)(n, m, dtype=dtype, device=device, transition_nnz_cap=M.nnz + n)
    try:
        w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(n, idx))
        st = w.walk(0.5, 80)
        return int(st.steps_done), w.iterate(), w.trace(), w
    finally:
        w.close()


def _trace_key(tr):
    return [(t["nnz_in"], t["nnz_new"], t["cut"] if t["step"] else 0.0, t["delta"]) for t in tr]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cached_structures_and_window_select_change_no_bit(cuda_device, dtype, monkeypatch):
    """The settled walk runs on the cached-structure kernel and the window select; with both switched off (row-group
    kernel and histogram select every step) the iterate, the cuts and the deltas are the same bits."""
    M, idx = synth.make_workload(synth.Workload("cs", 6000, 960, 100, 16))
    monkeypatch.setenv("CRWR_STRUCT", "2")          # rows this short do not get structures by default
    steps, it, tr, _ = _walk(M, idx, dtype, cuda_device)
    assert sum(t["fast_rows"] for t in tr[4:]) > 0.8 * idx.size * (len(tr) - 5)       # most rows, most steps
    assert all(tr[i]["fast_rows"] == 0 for i in range(4))
    monkeypatch.setenv("CRWR_STRUCT", "0")
    monkeypatch.setenv("CRWR_WIN_STEP", "0")
    steps0, it0, tr0, _ = _walk(M, idx, dtype, cuda_device)
    assert all(t["fast_rows"] == 0 for t in tr0)
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)


@pytest.mark.parametrize("env", [{"CRWR_POOL_BYTES": "300000"},                      # pool resets: structures rebuilt over and over
                                 {"CRWR_WIN_MIN": "1e-7", "CRWR_WIN_SCALE": "0.01"},   # windows that miss: full select in the tail
                                 {"CRWR_WIN_STEP": "2", "CRWR_WIN_REL_MAX": "10", "CRWR_WIN_MAX": "0.9"},   # window select from the first cut on
                                 {"CRWR_STRUCT_ALPHA": "1.0"},                        # K+ = K: structures fall out of date often
                                 {"CRWR_STRUCT_ALPHA": "0.05"},                       # very wide K+
                                 {"CRWR_WARP_SYM": "0"}])                             # first build by the CTA kernel (default for short rows: a warp per row)
def test_structure_and_window_corner_policies_change_no_bit(cuda_device, env, monkeypatch):
    M, idx = synth.make_workload(synth.Workload("cs", 5000, 800, 100, 16))
    steps0, it0, tr0, _ = _walk(M, idx, np.float32, cuda_device)
    monkeypatch.setenv("CRWR_STRUCT", "2")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    steps, it, tr, _ = _walk(M, idx, np.float32, cuda_device)
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_warp_built_structures_hand_long_rows_to_the_cta_kernel(cuda_device, dtype, monkeypatch):
    """Short rows get their structure from the warp-per-row kernel (structure_warp.cuh), whose tables follow the MEAN row;
    the rows of a few hub markers and of their neighbours, several times longer, exceed them and come back through the second list.  Same bits as the
    hash-accumulating kernels."""
    import scipy.sparse as sp
    M, idx = synth.make_workload(synth.Workload("hub", 6000, 900, 100, 16))
    rng = np.random.default_rng(7)
    n = M.shape[0]
    hubs = idx[rng.choice(idx.size, 3, replace=False)]
    rows = np.repeat(hubs, 300)
    cols = rng.integers(0, n, rows.size)
    keep = rows != cols
    add = sp.coo_matrix((rng.integers(1, 2, keep.sum()).astype(M.dtype), (rows[keep], cols[keep])), shape=M.shape).tocsr()
    add.sum_duplicates()
    M2 = (M + add + add.T).tocsr()
    M2.sort_indices()
    monkeypatch.setenv("CRWR_STRUCT", "0")
    steps0, it0, tr0, _ = _walk(M2, idx, dtype, cuda_device)
    assert max(t["max_row_len"] for t in tr0) > 3 * (tr0[3]["nnz_new"] // idx.size)       # well above the mean (the tables hold 2.2 x)
    monkeypatch.setenv("CRWR_STRUCT", "2")
    steps, it, tr, _ = _walk(M2, idx, dtype, cuda_device)
    assert sum(t["fast_rows"] for t in tr[4:]) > 0.7 * idx.size * (len(tr) - 5)
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)


@pytest.mark.parametrize("genome,degree", [(160, 20), (220, 24), (300, 30)])
def test_densities_between_the_bench_workloads_change_no_bit(cuda_device, genome, degree, monkeypatch):
    """Genome sizes between SURVEY 8d's two densities: the default policy (structures or not by the mean row, kept-list
    margin and overflow list by the row hints, window by the last two moves of the cut) against the plain path - row-group
    kernel and histogram select in every step.  No row may take the large-row kernel under the default policy."""
    M, idx = synth.make_workload(synth.Workload("mid", 12000, 1920, genome, degree))
    steps, it, tr, _ = _walk(M, idx, np.float32, cuda_device)
    assert all(t["fallback_rows"] == 0 for t in tr)
    monkeypatch.setenv("CRWR_STRUCT", "0")
    monkeypatch.setenv("CRWR_WIN_STEP", "0")
    steps0, it0, tr0, _ = _walk(M, idx, np.float32, cuda_device)
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)


def test_expanding_walk_keeps_the_histogram_select_and_changes_no_bit(cuda_device, monkeypatch):
    """rwr-thres 50: rows double per step and the cut falls several-fold - the host holds the window select for such a
    walk (crwr.expanding_walk -> crwr_hold_window_select).  Same bits as with the policy switched off."""
    from imputecc_b200 import crwr as crwr_mod
    M, idx = synth.make_workload(synth.Workload("exp", 6000, 960, 100, 16))
    seen = []
    real = crwr_mod.expanding_walk

    def spy(prev, now):
        seen.append(real(prev, now))
        return seen[-1]

    def walk50():
        n, m = M.shape[0], idx.size
        w = pkg.Walker(n, m, dtype=np.float32, device=cuda_device, transition_nnz_cap=M.nnz + n)
        try:
            w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(n, idx))
            st = w.walk(0.5, 50)
            return int(st.steps_done), w.iterate(), w.trace()
        finally:
            w.close()

    monkeypatch.setattr(crwr_mod, "expanding_walk", spy)
    steps, it, tr = walk50()
    assert seen and seen[-1] is True
    monkeypatch.setattr(crwr_mod, "expanding_walk", lambda prev, now: False)
    steps0, it0, tr0 = walk50()
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)


def test_deferred_rows_behind_a_window_step_are_retried(cuda_device):
    """Tiny accumulator hints from walk step 4 on: rows are deferred in a window-select step that queued no large-row
    kernel; the finish step flags it, the host walks again with the kernel queued, and the result is unchanged."""
    M, idx = synth.make_workload(synth.Workload("rt", 5000, 800, 100, 16))
    steps0, it0, tr0, _ = _walk(M, idx, np.float32, cuda_device)

    class LateTinyHints(pkg.Walker):
        def set_row_hints(self, max_row_len, max_kept):
            if max_row_len > 0 and getattr(self, "_calls", 0) >= 2 and self.retries == 0:      # the chunk from walk step 4 on
                max_row_len, max_kept = 24, 6
            self._calls = getattr(self, "_calls", 0) + 1
            super().set_row_hints(max_row_len, max_kept)

        def walk_init(self, *a, **k):
            self._calls = 0
            super().walk_init(*a, **k)

    steps, it, tr, w = _walk(M, idx, np.float32, cuda_device, walker_cls=LateTinyHints)
    assert w.retries == 1
    assert steps == steps0 and _trace_key(tr) == _trace_key(tr0)
    assert _bit_equal(it, it0)
