"""Subprocess body of tests/test_match_contigs.py: run match_contigs_b200 on one workload under the PYTHONHASHSEED the
golden vectors were produced with and print the result as JSON.  usage: match_worker.py TAG gpu|oracle"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from imputecc_b200.binning import match_contigs_b200, precluster_b200          # noqa: E402
from oracle.match_inputs import (SWEEP, SWEEP_WORKLOAD, WORKLOADS, OracleWeights, as_json, match_inputs,   # noqa: E402
                                 precluster_inputs)


def sweep(mode):
    """BASELINE config 5: every rp x thres cell, imputation and preliminary bins by the repo's chain."""
    impute = None
    if mode == "gpu":
        import imputecc_b200 as pkg
        impute = lambda M, idx, rp, thres: pkg.crwr_impute(M, idx, rp, thres, device="cuda:0")      # noqa: E731
    out = {}
    for rp, thres in SWEEP:
        E, rev, counts, marker_contigs, contig_markers = precluster_inputs(SWEEP_WORKLOAD, rp, thres, impute=impute)
        kw = {"_operator": OracleWeights(E)} if mode == "oracle" else {"device": "cuda:0"}
        bins, bin_of = precluster_b200(counts, marker_contigs, contig_markers, E, rev, 50, 0, **kw)
        out["rp%s_t%s" % (rp, thres)] = dict(bins={str(k): list(v) for k, v in bins.items()}, bin_of_contig=dict(bin_of),
                                             nnz=int(E.nnz))
    sys.stdout.write("RESULT " + json.dumps(out, sort_keys=True) + "\n")


def main():
    tag, mode = sys.argv[1], sys.argv[2]
    assert os.environ.get("PYTHONHASHSEED") == "0"
    if tag == "sweep":
        return sweep(mode)
    E, rev, contig_markers, smg, w_intra, w_inter = match_inputs(dict(WORKLOADS)[tag])
    kw = {"_operator": OracleWeights(E)} if mode == "oracle" else {"device": "cuda:0"}
    result = match_contigs_b200(smg, contig_markers, E.tolil(), rev, w_intra, w_inter, **kw)
    sys.stdout.write("RESULT " + json.dumps(as_json(result), sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
