"""match_contigs_b200 (imputecc_b200/binning.py) against the bins returned by the UNMODIFIED reference function
(tests/golden/match/*.json, produced by oracle/make_golden_match.py under PYTHONHASHSEED=0).  The reference's result
depends on the iteration order of sets of strings, so the function runs in a subprocess with the same hash seed."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO

TAGS = ["bw1500", "bw3000dense"]


def _run(tag, mode):
    env = dict(os.environ, PYTHONHASHSEED="0")
    out = subprocess.run([sys.executable, os.path.join(REPO, "tests", "match_worker.py"), tag, mode], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = next(l for l in out.stdout.splitlines() if l.startswith("RESULT "))
    return json.loads(line[len("RESULT "):])


def _golden(tag):
    with open(os.path.join(REPO, "tests", "golden", "match", "match_%s.json" % tag)) as fh:
        return json.load(fh)


@pytest.mark.parametrize("tag", TAGS)
def test_host_logic_matches_reference_bins(tag):
    """The iteration logic around the weights (sets, bipartite matching, bin updates) with the CPU oracle standing in for
    the GPU operator: identical bins, bin_of_contig, bin count, bin markers and binning order."""
    assert _run(tag, "oracle") == _golden(tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_path_matches_reference_bins(cuda_device, tag):
    """The drop-in as a user calls it: edge weights from the CUDA operator (99 iterations, matrix resident)."""
    got, want = _run(tag, "gpu"), _golden(tag)
    assert got["n_bins"] == want["n_bins"] and got["bins"] == want["bins"]
    assert got == want


def _sweep_golden():
    with open(os.path.join(REPO, "tests", "golden", "match", "precluster_sweep.json")) as fh:
        return json.load(fh)


def test_precluster_host_logic_matches_reference_over_the_sweep():
    """precluster_b200 (gene ordering, thresholds, empty-matrix branch, matching) on the oracle's imputed matrices with
    the oracle's weights: the bins of the unmodified PreCluster in all nine rp x thres cells."""
    assert _run("sweep", "oracle") == _sweep_golden()


@pytest.mark.gpu
def test_config5_sweep_preliminary_bins_identical(cuda_device):
    """BASELINE config 5: for rp in {0.3, 0.5, 0.7} x thres in {50, 80, 95} the repo's chain - crwr_impute on the GPU,
    preliminary bins with the edge weights from the GPU - returns exactly the bins that the reference's PreCluster
    derives from the reference's imputed matrix (thres 95 gives the empty matrix and the one-bin-per-contig branch)."""
    got, want = _run("sweep", "gpu"), _sweep_golden()
    assert sorted(got) == sorted(want)
    for cell in want:
        assert got[cell]["nnz"] == want[cell]["nnz"], cell
        assert got[cell]["bins"] == want[cell]["bins"] and got[cell]["bin_of_contig"] == want[cell]["bin_of_contig"], cell


def test_install_binning_rebinds_the_imported_name():
    import types
    import imputecc_b200.binning as binning
    pre, fin = types.ModuleType("pre_clustering"), types.ModuleType("final_cluster")
    pre.match_contigs = fin.match_contigs = lambda *a: None
    binning.install_binning(pre, fin)
    assert pre.match_contigs is binning.match_contigs_b200 and fin.match_contigs is binning.match_contigs_b200
