#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # needs /root/reference

Imports ``Script.utility.random_walk_cpu`` and ``Script.imputation.ImputeMatrix`` from
/root/reference and freezes, per case: the NormCC input (CSR triplets), the marker index
set, rp / perc, and the reference outputs
  * ``walk32``  random_walk_cpu(P float32)   - the reference-faithful path (utility.py:270)
  * ``walk64``  random_walk_cpu(P float64)   - same function, float64 P
  * ``E32``     ImputeMatrix(...).imputed_matrix driven through the real constructor with a
                fabricated HMMER --domtblout file (imputation.py:22-129) - bundled data only
together with the numpy / scipy versions that produced them.  P itself is produced by the
reference's own statements (imputation.py:112-119), executed here verbatim through
``ImputeMatrix._imputation`` for E32 and re-derived for the walk cases via the object.
The GPU box has no /root/reference; tests read only the .npz files written here.
"""
import os
import sys
import tempfile

import numpy as np
import scipy
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from Script import imputation as ref_imputation          # noqa: E402
from Script.utility import random_walk_cpu               # noqa: E402
from imputecc_b200 import synth                          # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def _csr_fields(prefix, M):
    M = sp.csr_matrix(M)
    M.sort_indices()
    return {prefix + "_indptr": M.indptr.astype(np.int64), prefix + "_indices": M.indices.astype(np.int32),
            prefix + "_data": M.data, prefix + "_shape": np.array(M.shape, dtype=np.int64)}


class _Probe(ref_imputation.ImputeMatrix):
    """ImputeMatrix whose inputs are injected instead of read from files.

    Only __init__ is replaced (file parsing is out of scope); ``_imputation`` is the
    reference's own method and runs unmodified.  ``walk_dtype`` lets the same statements
    feed random_walk_cpu a float64 P by wrapping the module-level name it calls.
    """

    def __init__(self, normcc, marker_idx, rp, perc, walk_dtype=None, capture=None):
        n = normcc.shape[0]
        names = np.array(["ctg%06d" % i for i in range(n)], dtype=object)
        self.contig_info = np.stack([names, np.zeros(n, dtype=object), np.ones(n, dtype=object)], axis=1)
        self.dict_contigRev = {nm: i for i, nm in enumerate(names)}
        self.normcc_matrix = normcc
        self.contig_markers = {names[i]: ["g"] for i in marker_idx}
        self.rwr_rp = rp
        self.rwr_perc = perc
        orig = ref_imputation.random_walk_cpu

        def spy(P, rp_, perc_, tol_, m_):
            if walk_dtype is not None:
                P = P.astype(walk_dtype)
            Q = orig(P, rp_, perc_, tol_, m_)
            if capture is not None:
                capture["P"] = P
                capture["Q"] = Q
            return Q

        ref_imputation.random_walk_cpu = spy
        try:
            self.contig_local, self.dict_contigRevLocal, self.imputed_matrix = self._imputation()
        finally:
            ref_imputation.random_walk_cpu = orig


def run_case(name, normcc, marker_idx, rp, perc):
    rec = {"rp": np.float64(rp), "perc": np.int64(perc), "marker_idx": np.asarray(marker_idx, dtype=np.int64),
           "numpy_version": np.array(np.__version__), "scipy_version": np.array(scipy.__version__)}
    rec.update(_csr_fields("normcc", normcc))
    for tag, dt in (("32", None), ("64", np.float64)):
        cap = {}
        obj = _Probe(normcc, marker_idx, rp, perc, walk_dtype=dt, capture=cap)
        rec.update(_csr_fields("walk" + tag, cap["Q"]))
        rec.update(_csr_fields("E" + tag, obj.imputed_matrix))
        if tag == "32":
            rec.update(_csr_fields("P32", cap["P"]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-28s m=%4d  walk32 nnz=%6d walk64 nnz=%6d" % (name, len(marker_idx), rec["walk32_data"].size,
                                                       rec["walk64_data"].size))


def bundled_constructor_case(normcc, info_csv, m, rp, perc):
    """Drive the real ImputeMatrix constructor with a fabricated hmmout."""
    import pandas as pd
    names = pd.read_csv(info_csv, sep=",", header=0).values[:, 0]
    idx = synth.top_degree_markers(normcc, m)
    with tempfile.TemporaryDirectory() as td:
        hmm = os.path.join(td, "contigs.hmmout")
        with open(hmm, "w") as fh:
            fh.write("# fabricated domtblout\n")
            for j, i in enumerate(idx):
                cols = ["%s_1_300_+" % names[i], "-", "100", "gene%03d" % (j % 107), "PF0", "100", "1e-30", "100", "0",
                        "1", "1", "1e-30", "1e-30", "100", "0", "1", "99", "1", "99", "1", "99", "0.99", "-"]
                fh.write(" ".join(cols) + "\n")
        mat = os.path.join(td, "m.npz")
        sp.save_npz(mat, normcc)
        obj = ref_imputation.ImputeMatrix(info_csv, mat, hmm, 0.9, rp, perc, 8000)
    return idx, obj


def main():
    os.makedirs(OUT, exist_ok=True)
    bundled = sp.load_npz("/root/reference/test_data/test_normalized_contact_matrix.npz")
    info_csv = "/root/reference/test_data/test_info.csv"
    # config 1: bundled matrix, deterministic top-degree marker sets
    for m in (60, 120, 250):
        for perc in (50, 80, 95):
            run_case("bundled_m%d_t%d_rp50" % (m, perc), bundled, synth.top_degree_markers(bundled, m), 0.5, perc)
    for rp in (0.3, 0.7):
        run_case("bundled_m120_t80_rp%d" % int(rp * 100), bundled, synth.top_degree_markers(bundled, 120), rp, 80)
    # the real constructor path, file parsing included, must agree with the injected one
    idx, obj = bundled_constructor_case(bundled, info_csv, 120, 0.5, 80)
    rec = {"marker_idx": idx.astype(np.int64)}
    rec.update(_csr_fields("E32", obj.imputed_matrix))
    rec["contig_local_names"] = np.array([str(x) for x in obj.contig_local[:, 0]])
    np.savez_compressed(os.path.join(OUT, "bundled_constructor_m120_t80_rp50.npz"), **rec)
    # random marker set on the bundled matrix: diagonal ties wipe the result
    rnd = np.sort(np.random.default_rng(7).choice(500, 100, replace=False))
    run_case("bundled_random100_t80_rp50", bundled, rnd, 0.5, 80)
    # small synthetic planted-partition cases, both densities, a few (rp, perc) cells
    small = synth.contact_matrix(1500, 50, 12, 2, seed=3)
    mk = synth.marker_set(1500, 240, seed=4)
    for rp, perc in ((0.5, 80), (0.3, 50), (0.7, 95), (0.5, 0), (0.5, 100)):
        run_case("synth1500_t%d_rp%d" % (perc, int(rp * 100)), small, mk, rp, perc)
    dense = synth.contact_matrix(3000, 300, 30, 2, seed=5)
    run_case("synth3000dense_t80_rp50", dense, synth.marker_set(3000, 480, seed=6), 0.5, 80)
    # asymmetric, diagonal-carrying input: exercises diag strip and column (not row) sums
    rng = np.random.default_rng(11)
    asym = sp.random(400, 400, density=0.02, random_state=rng, data_rvs=lambda k: rng.lognormal(3, 1, k)).tocsr()
    asym = asym + sp.diags(rng.random(400) * 5)
    run_case("asym400_t60_rp50", asym.tocsr(), synth.marker_set(400, 80, seed=12), 0.5, 60)


if __name__ == "__main__":
    main()
