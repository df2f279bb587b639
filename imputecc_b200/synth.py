"""Synthetic NormCC-like contact matrices.

A planted-partition graph: contigs belong to genomes; most Hi-C contacts are
intra-genome with log-normal weights, a few are spurious inter-genome contacts
with weaker weights.  The result looks like the bundled
``test_normalized_contact_matrix.npz``: symmetric, zero diagonal, float64
values, sorted int32 indices.  Used by ``bench.py`` and the tests; it is input
generation only, not part of the imputation path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass(frozen=True)
class Workload:
    name: str
    n_contigs: int
    n_markers: int
    genome_size: int
    intra_degree: int
    inter_degree: int = 2
    matrix_seed: int = 0
    marker_seed: int = 1


# BASELINE.json configs[1..3] at the two densities SURVEY.md §8d defines.
WORKLOADS = {
    "c2_sparse": Workload("c2_sparse", 12_500, 2_000, 100, 16),
    "c2_dense": Workload("c2_dense", 12_500, 2_000, 500, 40),
    "c3_sparse": Workload("c3_sparse", 50_000, 8_000, 100, 16),
    "c3_dense": Workload("c3_dense", 50_000, 8_000, 500, 40),
    "c4_sparse": Workload("c4_sparse", 200_000, 32_000, 100, 16),
    "c4_dense": Workload("c4_dense", 200_000, 32_000, 500, 40),
}


def contact_matrix(n_contigs: int, genome_size: int, intra_degree: int, inter_degree: int = 2,
                   seed: int = 0) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    n = n_contigs
    n_genomes = max(n // genome_size, 1)
    genome_of = np.sort(rng.integers(0, n_genomes, n))
    first = np.searchsorted(genome_of, np.arange(n_genomes), side="left")
    last = np.searchsorted(genome_of, np.arange(n_genomes), side="right")

    n_intra = n * intra_degree // 2
    u = rng.integers(0, n, n_intra)
    g = genome_of[u]
    v = first[g] + np.floor(rng.random(n_intra) * (last[g] - first[g])).astype(np.int64)
    w = rng.lognormal(6.2, 1.12, n_intra)

    n_inter = n * inter_degree // 2
    u2 = rng.integers(0, n, n_inter)
    v2 = rng.integers(0, n, n_inter)
    w2 = rng.lognormal(4.2, 1.12, n_inter)

    rows = np.concatenate([u, u2])
    cols = np.concatenate([v, v2])
    vals = np.concatenate([w, w2])
    keep = rows != cols
    # A = coo -> csr (duplicates summed); A = triu(A, 1); A = A + A^T.  Pairs drawn
    # with u > v fall below the diagonal and are discarded by triu, as in the survey's generator.
    upper = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n)).tocsr()
    upper = sp.triu(upper, 1).tocsr()
    full = (upper + upper.T).tocsr()
    full.sort_indices()
    full.indices = full.indices.astype(np.int32)
    full.indptr = full.indptr.astype(np.int32)
    return full


def marker_set(n_contigs: int, n_markers: int, seed: int = 1) -> np.ndarray:
    return np.sort(np.random.default_rng(seed).choice(n_contigs, n_markers, replace=False)).astype(np.int64)


def make_workload(w: Workload):
    """(normcc CSR float64, sorted marker indices) for a named workload."""
    return (contact_matrix(w.n_contigs, w.genome_size, w.intra_degree, w.inter_degree, w.matrix_seed),
            marker_set(w.n_contigs, w.n_markers, w.marker_seed))


def top_degree_markers(normcc: sp.csr_matrix, n_markers: int) -> np.ndarray:
    """Deterministic marker set for the bundled test matrix."""
    deg = np.diff(normcc.tocsr().indptr)
    return np.sort(np.argsort(-deg, kind="stable")[:n_markers]).astype(np.int64)
