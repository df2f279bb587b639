"""Inputs of ``match_contigs`` for vectors and tests.  TEST INFRASTRUCTURE ONLY.

Nothing here needs the reference: planted single-copy genes, the iteration order PreCluster derives from them, the
imputed matrix (from the CPU oracle, bit-identical to the reference's) and the thresholds ImputeCC passes.
"""
import numpy as np
import scipy.sparse as sp

from imputecc_b200 import synth
from oracle import binweights_oracle as bw
from oracle import crwr_oracle as oracle

WORKLOADS = (("bw1500", synth.Workload("bw1500", 1500, 240, 100, 16)),
             ("bw3000dense", synth.Workload("bw3000dense", 3000, 300, 300, 30)))


def planted_markers(wl, idx, rng):
    """Each marker contig carries 1-3 of 107 single-copy genes, no gene twice inside a planted genome
    (SURVEY.md §8d config 5)."""
    n = wl.n_contigs
    g = n // wl.genome_size
    lab = np.sort(np.random.default_rng(wl.matrix_seed).integers(0, g, n))      # the generator's genome labels
    used = {}
    contig_markers = {}
    for i in idx:
        k = int(rng.integers(1, 4))
        free = [x for x in range(107) if x not in used.setdefault(int(lab[i]), set())]
        genes = [int(x) for x in rng.choice(free, size=min(k, len(free)), replace=False)]
        used[int(lab[i])].update(genes)
        contig_markers["contig%05d" % i] = ["gene%03d" % x for x in genes]
    return contig_markers


def smg_iterations(contig_markers):
    """The iteration order PreCluster derives (pre_clustering.py:11-44): genes by contig count, ties by the total
    number of genes on their contigs - restated for input generation only."""
    marker_contigs = {}
    for c, genes in contig_markers.items():
        for gname in genes:
            marker_contigs.setdefault(gname, []).append(c)
    counts = {gname: len(cs) for gname, cs in marker_contigs.items()}
    out, n = {}, 0
    for cnt in sorted(set(counts.values()), reverse=True):
        tot = {gname: sum(len(contig_markers[c]) for c in marker_contigs[gname]) for gname in counts if counts[gname] == cnt}
        for gname, _ in sorted(tot.items(), key=lambda kv: kv[1], reverse=True):
            out[n] = marker_contigs[gname]
            n += 1
    return out


class OracleWeights:
    """Same call as imputecc_b200.BinContactWeights.weights, computed by the CPU oracle."""

    def __init__(self, E):
        self.E = sp.csr_matrix(E)

    def weights(self, to_bin, bin_ptr, bin_members):
        return bw.bin_contact_weights(self.E, to_bin, bin_ptr, bin_members)


def match_inputs(wl):
    """Imputed matrix, name -> row, marker genes, iteration order, w_intra, w_inter (final_cluster.py:401-408 style)."""
    M, idx = synth.make_workload(wl)
    E = sp.csr_matrix(oracle.crwr_impute(M, idx, 0.5, 80, dtype=np.float32))
    E.sort_indices()
    names = ["contig%05d" % i for i in idx]
    rev = {nm: k for k, nm in enumerate(names)}
    contig_markers = planted_markers(wl, idx, np.random.default_rng(2))
    smg = smg_iterations(contig_markers)
    data = E.tocoo().data
    return E, rev, contig_markers, smg, np.percentile(data, 50), np.percentile(data, 0)


def as_json(result):
    bins, bin_of_contig, n_bins, bin_markers, binned = result
    return dict(bins={str(k): list(v) for k, v in bins.items()}, bin_of_contig=dict(bin_of_contig), n_bins=int(n_bins),
                bin_markers={str(k): list(v) for k, v in bin_markers.items()}, binned=list(binned))


# ---- BASELINE config 5: the rp x thres sweep with a preliminary-bin identity check
SWEEP = [(rp, thres) for rp in (0.3, 0.5, 0.7) for thres in (50, 80, 95)]
SWEEP_WORKLOAD = synth.Workload("sweep2000", 2000, 320, 100, 16)


def marker_tables(wl, idx):
    """(marker_contig_counts, marker_contigs, contig_markers) as utility.py:331-348 derives them from the HMMER table."""
    contig_markers = planted_markers(wl, idx, np.random.default_rng(2))
    marker_contigs = {}
    for c, genes in contig_markers.items():
        for g in genes:
            marker_contigs.setdefault(g, []).append(c)
    counts = {g: len(cs) for g, cs in marker_contigs.items()}
    return counts, marker_contigs, contig_markers


def precluster_inputs(wl, rp, thres, impute=None):
    """Imputed matrix (``impute(M, idx, rp, thres)``; default: the CPU oracle), name -> row, and the marker tables."""
    M, idx = synth.make_workload(wl)
    if impute is None:
        E = sp.csr_matrix(oracle.crwr_impute(M, idx, rp, thres, dtype=np.float32))
    else:
        E = sp.csr_matrix(impute(M, idx, rp, thres))
    E.sort_indices()
    rev = {"contig%05d" % i: k for k, i in enumerate(idx)}
    counts, marker_contigs, contig_markers = marker_tables(wl, idx)
    return E, rev, counts, marker_contigs, contig_markers
