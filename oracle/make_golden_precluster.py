"""Golden preliminary bins (BASELINE config 5: parameter sweep with a bin identity check).

TEST INFRASTRUCTURE ONLY.  For every cell of the rp x thres sweep on a planted-partition workload: the UNMODIFIED
reference imputation statements (through the oracle, which tests/test_oracle.py pins bit for bit to the reference's
own output) feed the UNMODIFIED ``Script.pre_clustering.PreCluster`` (/root/reference/Script/pre_clustering.py:10-63),
and the bins it returns are frozen.  tests/test_match_contigs.py then runs the repo's chain - crwr_impute on the GPU,
precluster_b200 with the edge weights from the GPU - and demands the same bins.  PYTHONHASHSEED=0 throughout (the
reference iterates sets of strings); the script re-executes itself with it.  It also checks here that the host mirror
``precluster_b200``, with the CPU oracle standing in for the GPU weights, returns the identical bins.

    PYTHONPATH=/root/reference python oracle/make_golden_precluster.py
"""
import json
import os
import sys
import types

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
for name in ("igraph", "leidenalg", "Bio", "Bio.SeqIO"):     # imported at module top by the reference, unused on this path
    sys.modules.setdefault(name, types.ModuleType(name))

from Script import pre_clustering as ref_pre                   # noqa: E402  (needs PYTHONPATH=/root/reference)
from imputecc_b200.binning import precluster_b200              # noqa: E402
from oracle.match_inputs import OracleWeights, precluster_inputs, SWEEP, SWEEP_WORKLOAD   # noqa: E402


def main():
    out_dir = os.path.join(REPO, "tests", "golden", "match")
    os.makedirs(out_dir, exist_ok=True)
    out = {}
    for rp, thres in SWEEP:
        E, rev, counts, marker_contigs, contig_markers = precluster_inputs(SWEEP_WORKLOAD, rp, thres)
        bins, bin_of = ref_pre.PreCluster(counts, marker_contigs, contig_markers, E.tolil(), rev, 50, 0)
        got_bins, got_of = precluster_b200(counts, marker_contigs, contig_markers, E, rev, 50, 0, _operator=OracleWeights(E))
        assert got_bins == bins and got_of == bin_of, "the host mirror diverges from PreCluster at rp %s thres %s" % (rp, thres)
        out["rp%s_t%s" % (rp, thres)] = dict(bins={str(k): list(v) for k, v in bins.items()}, bin_of_contig=dict(bin_of),
                                             nnz=int(E.nnz))
        print("rp", rp, "thres", thres, "imputed nnz", E.nnz, "bins", len(bins), "- mirror identical")
    with open(os.path.join(out_dir, "precluster_sweep.json"), "w") as fh:
        json.dump(out, fh, sort_keys=True)


if __name__ == "__main__":
    main()
