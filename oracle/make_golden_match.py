"""Golden results of ``match_contigs``: the bins themselves.

TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference function ``Script.utility.match_contigs``
(/root/reference/Script/utility.py:409-590) in the build container on the two workloads of
oracle/make_golden_binweights.py and freezes what it returns (bins, bin_of_contig, n_bins, bin_markers, the list
of binned contigs).  The reference iterates sets and dicts of strings, so its result depends on PYTHONHASHSEED:
vectors and tests run under PYTHONHASHSEED=0 (this script re-executes itself with it).  It also checks here, once,
that the repo's host mirror ``match_contigs_b200`` - with the CPU oracle standing in for the GPU weights, which are the
same bits (tests/test_binweights.py) - returns the identical tuple.

    PYTHONPATH=/root/reference python oracle/make_golden_match.py
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

from Script import utility as ref_utility                      # noqa: E402  (needs PYTHONPATH=/root/reference)
from imputecc_b200.binning import match_contigs_b200           # noqa: E402
from oracle.match_inputs import WORKLOADS, OracleWeights, as_json, match_inputs   # noqa: E402


def main():
    out_dir = os.path.join(REPO, "tests", "golden", "match")
    os.makedirs(out_dir, exist_ok=True)
    for tag, wl in WORKLOADS:
        E, rev, contig_markers, smg, w_intra, w_inter = match_inputs(wl)
        want = as_json(ref_utility.match_contigs(smg, contig_markers, E.tolil(), rev, w_intra, w_inter))
        got = as_json(match_contigs_b200(smg, contig_markers, E.tolil(), rev, w_intra, w_inter, _operator=OracleWeights(E)))
        assert got == want, "the host mirror diverges from the reference on %s" % tag
        with open(os.path.join(out_dir, "match_%s.json" % tag), "w") as fh:
            json.dump(want, fh, sort_keys=True)
        print(tag, "bins", want["n_bins"], "binned", len(want["binned"]), "- mirror identical")


if __name__ == "__main__":
    main()
