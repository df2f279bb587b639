"""Bin-contact weights of match_contigs: oracle vs the vectors recorded from the unmodified
reference function, and (gpu) the CUDA operator vs the oracle."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import binweights_oracle as bw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "binweights")
CASES = ["binweights_bw1500", "binweights_bw3000dense"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    m = int(z["m"])
    E = sp.csr_matrix((z["E_data"], z["E_indices"], z["E_indptr"]), shape=(m, m))
    recs = [dict(bin_ptr=z["r%d_bin_ptr" % r], bin_members=z["r%d_bin_members" % r], to_bin=z["r%d_to_bin" % r],
                 weights=z["r%d_weights" % r]) for r in range(int(z["n_records"]))]
    return E, recs


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_vectors(name):
    E, recs = load(name)
    assert E.dtype == np.float32 and len(recs) >= 3
    for rec in recs:
        got = bw.bin_contact_weights(E, rec["to_bin"], rec["bin_ptr"], rec["bin_members"])
        assert got.dtype == np.float32 and got.shape == rec["weights"].shape
        assert np.array_equal(bits(got), bits(rec["weights"]))                 # bit for bit, -0.0 included
    rec = recs[0]
    loops = bw.bin_contact_weights_loops(E, rec["to_bin"], rec["bin_ptr"], rec["bin_members"])
    assert np.array_equal(bits(loops.astype(np.float32)), bits(rec["weights"]))


def test_bins_to_arrays_follows_bin_index_and_list_order():
    from imputecc_b200 import bins_to_arrays
    ptr, mem = bins_to_arrays({2: ["c", "a"], 0: ["b"]}, {"a": 0, "b": 1, "c": 2})
    assert ptr.tolist() == [0, 1, 3] and mem.tolist() == [1, 2, 0]


def test_empty_contact_gives_negative_zero():
    E = sp.csr_matrix((4, 4), dtype=np.float32)
    got = bw.bin_contact_weights(E, np.array([3]), np.array([0, 2, 3]), np.array([0, 1, 2], dtype=np.int32))
    assert np.array_equal(bits(got), bits(np.array([[-0.0, -0.0]], dtype=np.float32)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_vectors(cuda_device, name):
    import imputecc_b200 as pkg
    E, recs = load(name)
    resident = pkg.BinContactWeights(E, device=cuda_device)          # the matrix stays on the GPU across iterations
    for rec in recs:
        got = pkg.bin_contact_weights(E, rec["to_bin"], rec["bin_ptr"], rec["bin_members"], device=cuda_device)
        assert got.dtype == np.float32 and np.array_equal(bits(got), bits(rec["weights"]))
        again = resident.weights(rec["to_bin"], rec["bin_ptr"], rec["bin_members"])
        assert np.array_equal(bits(again), bits(rec["weights"]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_matches_oracle_on_random_bins(cuda_device, dtype):
    import imputecc_b200 as pkg
    rng = np.random.default_rng(5)
    m = 3000
    E = sp.random(m, m, density=0.01, format="csr", dtype=np.float64, random_state=7).astype(dtype)
    E.sort_indices()
    perm = rng.permutation(m)
    to_bin = np.sort(perm[:300]).astype(np.int32)
    rest = perm[300:2500]
    cuts = np.sort(rng.choice(np.arange(1, rest.size), size=199, replace=False))
    bin_ptr = np.concatenate([[0], cuts, [rest.size]]).astype(np.int64)
    got = pkg.bin_contact_weights(E, to_bin, bin_ptr, rest.astype(np.int32), device=cuda_device)
    want = bw.bin_contact_weights(E, to_bin, bin_ptr, rest.astype(np.int32))
    assert got.dtype == dtype and np.array_equal(bits(got), bits(want))


@pytest.mark.gpu
def test_gpu_long_rows_and_overlapping_bins(cuda_device):
    """Rows longer than the shared-memory list take the binary-search form inside the row kernel; bins that share a
    contig (never produced by the reference, but legal input) take the pair kernel."""
    import imputecc_b200 as pkg
    rng = np.random.default_rng(9)
    m = 1500
    E = sp.random(m, m, density=0.25, format="csr", dtype=np.float32, random_state=3)     # ~375 stored entries per row
    E.sort_indices()
    assert np.diff(E.indptr).max() > 256
    perm = rng.permutation(m)
    to_bin = np.sort(perm[:64]).astype(np.int32)
    rest = perm[64:1200].astype(np.int32)
    bin_ptr = np.arange(0, rest.size + 1, 8, dtype=np.int64)
    got = pkg.bin_contact_weights(E, to_bin, bin_ptr, rest, device=cuda_device)
    assert np.array_equal(bits(got), bits(bw.bin_contact_weights(E, to_bin, bin_ptr, rest)))
    overlapping = np.concatenate([rest[:40], rest[:40]])                                  # every contig in two bins
    ptr2 = np.arange(0, overlapping.size + 1, 10, dtype=np.int64)
    got = pkg.bin_contact_weights(E, to_bin, ptr2, overlapping, device=cuda_device)
    assert np.array_equal(bits(got), bits(bw.bin_contact_weights(E, to_bin, ptr2, overlapping)))


@pytest.mark.gpu
def test_gpu_argument_errors(cuda_device):
    import imputecc_b200 as pkg
    E = sp.identity(5, dtype=np.float32, format="csr")
    with pytest.raises(ValueError):
        pkg.bin_contact_weights(E, np.array([7]), np.array([0, 1]), np.array([0], dtype=np.int32), device=cuda_device)
    with pytest.raises(ValueError):
        pkg.bin_contact_weights(E, np.array([1]), np.array([0, 0]), np.array([], dtype=np.int32), device=cuda_device)   # empty bin
    out = pkg.bin_contact_weights(E, np.array([], dtype=np.int32), np.array([0, 1]), np.array([0], dtype=np.int32), device=cuda_device)
    assert out.shape == (0, 1)
