"""Golden vectors for the bin-contact weights of ``match_contigs``.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference function
``Script.utility.match_contigs`` (/root/reference/Script/utility.py:409-590) in the build
container and records, with a trace hook, what its edge-weight loop (utility.py:466-480)
computed in every iteration: the bins as they were at that moment, the contigs to bin,
and the weight of every (bin, contig) edge.  Nothing of the reference is copied; the
function is executed as shipped.  The imputed matrix fed to it is the oracle's
(bit-identical to the reference's, tests/test_oracle.py).

    PYTHONPATH=/root/reference PYTHONHASHSEED=0 python oracle/make_golden_binweights.py
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
for name in ("igraph", "leidenalg", "Bio", "Bio.SeqIO"):     # imported at module top by the reference, unused on this path
    sys.modules.setdefault(name, types.ModuleType(name))

from Script import utility as ref_utility          # noqa: E402  (needs PYTHONPATH=/root/reference)
from imputecc_b200 import synth                    # noqa: E402
from oracle import crwr_oracle as oracle           # noqa: E402
from oracle.match_inputs import planted_markers, smg_iterations   # noqa: E402,F401


def trace_match_contigs(smg_iteration, contig_markers, lil, rev, w_intra, w_inter):
    """Run the reference function; every time its edge list has just been completed, copy the locals."""
    records = []
    code = ref_utility.match_contigs.__code__

    def tracer(frame, event, arg):
        if frame.f_code is not code:
            return None
        def local(frame, event, arg):
            if event == "line":
                loc = frame.f_locals
                # the statement right after the edge loop is `B.add_nodes_from(top_nodes, bipartite=0)`
                if "edges" in loc and loc.get("edges") and loc.get("_rec_len") != len(loc["edges"]):
                    src = frame.f_code.co_filename
                    import linecache
                    if "add_nodes_from(top_nodes" in linecache.getline(src, frame.f_lineno):
                        records.append(dict(bins={k: list(v) for k, v in loc["bins"].items()}, to_bin=list(loc["to_bin"]),
                                            edges=list(loc["edges"])))
            return local
        return local

    sys.settrace(tracer)
    try:
        result = ref_utility.match_contigs(smg_iteration, contig_markers, lil, rev, w_intra, w_inter)
    finally:
        sys.settrace(None)
    return result, records


def main():
    out_dir = os.path.join(REPO, "tests", "golden", "binweights")
    for tag, wl in (("bw1500", synth.Workload("bw1500", 1500, 240, 100, 16)),
                    ("bw3000dense", synth.Workload("bw3000dense", 3000, 300, 300, 30))):
        M, idx = synth.make_workload(wl)
        E = sp.csr_matrix(oracle.crwr_impute(M, idx, 0.5, 80, dtype=np.float32))
        E.sort_indices()
        names = ["contig%05d" % i for i in idx]
        rev = {nm: k for k, nm in enumerate(names)}
        contig_markers = planted_markers(wl, idx, np.random.default_rng(2))
        smg = smg_iterations(contig_markers)
        data = E.tocoo().data
        (bins, bin_of, _, _, _), recs = trace_match_contigs(smg, contig_markers, E.tolil(), rev, np.percentile(data, 50),
                                                            np.percentile(data, 0))
        assert recs, "the trace hook saw no iteration"
        big = max(range(len(recs)), key=lambda i: len(recs[i]["to_bin"]) * len(recs[i]["bins"]))
        keep = [recs[i] for i in sorted({0, big, len(recs) // 2, len(recs) - 1})]     # first, largest, middle, last iteration
        arrays = dict(E_indptr=E.indptr.astype(np.int64), E_indices=E.indices.astype(np.int32), E_data=E.data.astype(np.float32),
                      m=np.int64(E.shape[0]), n_records=np.int64(len(keep)), n_iterations=np.int64(len(recs)),
                      final_bins=np.int64(len(bins)))
        for r, rec in enumerate(keep):
            keys = sorted(rec["bins"])
            members = [np.array([rev[c] for c in rec["bins"][k]], dtype=np.int32) for k in keys]
            rep = {rec["bins"][k][0]: b for b, k in enumerate(keys)}          # an edge names a bin by its first contig
            to_bin = [rev[c] for c in rec["to_bin"]]
            W = np.full((len(to_bin), len(keys)), np.nan, dtype=np.float32)
            row_of = {c: i for i, c in enumerate(rec["to_bin"])}
            for first, contig, w in rec["edges"]:
                W[row_of[contig], rep[first]] = np.float32(w)
                assert isinstance(w, np.floating) and np.dtype(type(w)) == np.float32, type(w)
            assert not np.isnan(W).any()
            arrays["r%d_bin_ptr" % r] = np.concatenate([[0], np.cumsum([len(x) for x in members])]).astype(np.int64)
            arrays["r%d_bin_members" % r] = np.concatenate(members).astype(np.int32)
            arrays["r%d_to_bin" % r] = np.array(to_bin, dtype=np.int32)
            arrays["r%d_weights" % r] = W
        np.savez_compressed(os.path.join(out_dir, "binweights_%s.npz" % tag), **arrays)
        print(tag, "iterations", len(recs), "kept", [(len(k["to_bin"]), len(k["bins"])) for k in keep], "final bins", len(bins))


if __name__ == "__main__":
    main()
