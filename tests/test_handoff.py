"""The hand-off of the imputed matrix to the clustering stages (imputecc_b200/handoff.py)."""
import gzip
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import scipy.sparse as sp

import imputecc_b200 as pkg
from conftest import load_csr, load_golden


def _imputed():
    return load_csr(load_golden("bundled_m120_t50_rp50"), "E32")          # an imputed matrix as the reference produced it


@pytest.mark.parametrize("fmt", ["lil", "csr", "csc", "coo"])
def test_contact_thresholds_equal_the_reference_expressions(fmt):
    E = _imputed()
    M = getattr(E, "to" + fmt)()
    want = (np.percentile(M.tocoo().data, 80), np.percentile(M.tocoo().data, 10))       # pre_clustering.py:60-61
    got = pkg.contact_thresholds(M, 80, 10)
    assert got[0] == want[0] and got[1] == want[1] and type(got[0]) is type(want[0])
    assert pkg.contact_thresholds(sp.lil_matrix((5, 5), dtype=np.float32), 80, 10) is None     # pre_clustering.py:47


def test_storage_file_loads_without_this_package(tmp_path):
    """ImputeCC.py:241 pickles the ImputeMatrix object (utility.py:26-35, gzip), ImputeCC.py:338 unpickles it in
    another process: what the GPU path puts into it must be plain numpy / scipy."""
    E = _imputed()
    names = np.array(["ctg%03d" % i for i in range(E.shape[0])], dtype=object)
    members = dict(contig_local=np.stack([names, names, names], axis=1), dict_contigRevLocal={n: i for i, n in enumerate(names)},
                   imputed_matrix=E.tolil(), imputed_csr=E)                # both hand-off forms of imputation_b200
    assert pkg.storage_offenders(members) == []
    path = tmp_path / "ImputeCC_storage.gz"
    with gzip.open(path, "wb") as fh:
        pickle.dump(members, fh)
    code = textwrap.dedent("""
        import gzip, pickle, sys
        sys.modules["imputecc_b200"] = None            # importing the package would raise: the file must not need it
        sys.modules["torch"] = None
        with gzip.open(sys.argv[1], "rb") as fh:
            obj = pickle.load(fh)
        lil, csr = obj["imputed_matrix"], obj["imputed_csr"]
        assert type(lil).__name__ == "lil_matrix" and (lil.tocsr() != csr).nnz == 0
        print("OK", csr.nnz, len(obj["dict_contigRevLocal"]))
    """)
    out = subprocess.run([sys.executable, "-c", code, str(path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("OK %d %d" % (E.nnz, E.shape[0])), out.stderr[-1000:]


def test_storage_offenders_names_gpu_bound_members():
    import ctypes
    import torch

    class Holder:
        pass

    h = Holder()
    h.matrix = _imputed()
    h.extra = {"t": torch.zeros(2), "ok": [1, 2.0, "x", np.float32(1)]}
    h.handle = ctypes.c_void_p(0)
    bad = pkg.storage_offenders(h)
    assert len(bad) == 2 and any("torch" in b for b in bad) and any("ctypes" in b for b in bad)
