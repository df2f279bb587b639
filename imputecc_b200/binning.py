"""Bin-contact weights of ``match_contigs`` on the GPU (the step after the imputation).

The reference (/root/reference/Script/utility.py:466-480) computes, in every iteration of its matching, the mean
imputed contact between each contig to bin and each existing bin with scalar look-ups in a LIL matrix - measured
6 s at 2,000 markers against 0.4 s for the walk itself.  :func:`bin_contact_weights` returns the same weights
(bit for bit: same float sums in the same order, same ``-0.0`` for a bin without contact) from one kernel launch;
the bipartite matching on those weights stays on the CPU, as in the reference.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import CRWR_F32, CRWR_F64
from .crwr import _require_cuda, _torch


def bins_to_arrays(bins, index_of=None):
    """``bins``: the reference's ``{bin index: [contig, ...]}`` dict (or a list of lists) -> (bin_ptr int64, members
    int32) in ascending bin-index order, members in list order.  ``index_of`` maps contig names to local indices
    (``dict_contigRevLocal``) when the bins hold names."""
    keys = sorted(bins) if isinstance(bins, dict) else range(len(bins))
    lists = [bins[k] for k in keys]
    if index_of is not None:
        lists = [[index_of[c] for c in lst] for lst in lists]
    ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(lst) for lst in lists])
    members = np.fromiter((c for lst in lists for c in lst), dtype=np.int32, count=int(ptr[-1]))
    return ptr, members


class BinContactWeights:
    """The imputed matrix kept on the GPU for the ~100 iterations of one ``match_contigs`` call: upload once, then one
    small launch per iteration.  ``weights(to_bin, bin_ptr, bin_members)`` as :func:`bin_contact_weights`."""

    def __init__(self, E, device=None):
        torch = _torch()
        self.device = _require_cuda(device)
        E = sp.csr_matrix(E)
        if E.shape[0] != E.shape[1]:
            raise ValueError("E must be square")
        if E.dtype not in (np.float32, np.float64):
            raise ValueError("E must be float32 or float64")
        if not E.has_sorted_indices:
            E = E.copy()
            E.sort_indices()
        self.m = E.shape[0]
        self.np_dtype = E.dtype
        self.code = CRWR_F64 if E.dtype == np.float64 else CRWR_F32
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            self.d_ip = self._up(E.indptr.astype(np.int64, copy=False))
            self.d_ix = self._up(E.indices.astype(np.int32, copy=False) if E.nnz else np.zeros(1, np.int32))
            self.d_dt = self._up(E.data if E.nnz else np.zeros(1, E.dtype))
            self.scratch_bytes = int(self.lib.crwr_bin_contact_scratch_bytes(self.m))
            self.d_scratch = torch.empty(self.scratch_bytes, dtype=torch.uint8, device=self.device)   # inverse map of the bins

    def _up(self, a):
        return _torch().from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=True)

    def weights(self, to_bin, bin_ptr, bin_members) -> np.ndarray:
        torch = _torch()
        m = self.m
        to_bin = np.ascontiguousarray(to_bin, dtype=np.int32)
        bin_ptr = np.ascontiguousarray(bin_ptr, dtype=np.int64)
        bin_members = np.ascontiguousarray(bin_members, dtype=np.int32)
        n_bins = bin_ptr.size - 1
        if n_bins < 0 or (bin_ptr.size and bin_ptr[0] != 0) or (n_bins >= 0 and bin_ptr[-1] != bin_members.size):
            raise ValueError("bin_ptr must start at 0 and end at len(bin_members)")
        if n_bins > 0 and np.any(np.diff(bin_ptr) <= 0):
            raise ValueError("every bin must hold at least one contig")
        for name, arr in (("to_bin", to_bin), ("bin_members", bin_members)):
            if arr.size and (arr.min() < 0 or arr.max() >= m):
                raise ValueError("%s holds an index outside [0, %d)" % (name, m))
        disjoint = int(np.bincount(bin_members, minlength=1).max()) <= 1      # as in the reference: a contig sits in one bin
        out_shape = (to_bin.size, max(n_bins, 0))
        if to_bin.size == 0 or n_bins <= 0:
            return np.zeros(out_shape, dtype=self.np_dtype)
        with torch.cuda.device(self.device):
            d_rows, d_ptr, d_mem = self._up(to_bin), self._up(bin_ptr), self._up(bin_members)
            d_out = torch.empty(out_shape, dtype=torch.float64 if self.code == CRWR_F64 else torch.float32, device=self.device)
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.lib.crwr_bin_contact_weights(self.code, m, C.c_void_p(self.d_ip.data_ptr()),
                                                         C.c_void_p(self.d_ix.data_ptr()), C.c_void_p(self.d_dt.data_ptr()),
                                                         to_bin.size, C.c_void_p(d_rows.data_ptr()), n_bins,
                                                         C.c_void_p(d_ptr.data_ptr()), C.c_void_p(d_mem.data_ptr()), bin_members.size,
                                                         C.c_void_p(self.d_scratch.data_ptr()) if disjoint else None,
                                                         self.scratch_bytes if disjoint else 0,
                                                         C.c_void_p(d_out.data_ptr()), stream))
            return d_out.cpu().numpy()


def bin_contact_weights(E, to_bin, bin_ptr, bin_members, device=None) -> np.ndarray:
    """weights[r, b] = -(sum_j E[to_bin[r], members_b[j]]) / len(members_b), utility.py:466-480.

    ``E``: imputed m x m scipy matrix (float32 as the reference's, or float64); ``to_bin``: local contig indices;
    ``bin_ptr`` / ``bin_members``: the bins in CSR form (see :func:`bins_to_arrays`), every bin non-empty.
    Returns a ``(len(to_bin), n_bins)`` array of E's dtype.  One-shot form of :class:`BinContactWeights`.
    """
    return BinContactWeights(E, device).weights(to_bin, bin_ptr, bin_members)


# ---------------------------------------------------------------------------------------------
# match_contigs with the edge weights from the GPU (drop-in for Script.utility.match_contigs)
# ---------------------------------------------------------------------------------------------

def _seed_bins(first_contigs, contig_markers):
    """utility.py:431-441: every contig carrying the first single-copy gene opens a bin of its own."""
    bins, bin_of_contig, bin_markers = {}, {}, {}
    for b, contig in enumerate(first_contigs):
        bins[b] = [contig]
        bin_of_contig[contig] = b
        bin_markers[b] = contig_markers[contig]
    return bins, bin_of_contig, bin_markers


def match_contigs_b200(smg_iteration, contig_markers, normcc_matrix, dict_contigRev, w_intra, w_inter, *, device=None,
                       _operator=None):
    """Drop-in for ``match_contigs(smg_iteration, contig_markers, normcc_matrix, dict_contigRev, w_intra, w_inter)``
    (/root/reference/Script/utility.py:409-590; called at pre_clustering.py:55 and final_cluster.py:401).

    Same arguments - ``smg_iteration``: {iteration: [contig names carrying that single-copy gene]}, ``contig_markers``:
    {contig: [genes]}, ``normcc_matrix``: the imputed matrix (the reference passes a LIL matrix; any scipy sparse
    matrix is taken), ``dict_contigRev``: contig name -> row of that matrix, ``w_intra`` / ``w_inter``: the two
    contact thresholds - and the same 5-tuple back: ``(bins, bin_of_contig, n_bins, bin_markers,
    binned_contigs_with_markers)``.

    What moves to the GPU is the loop nest of utility.py:466-480 - the mean imputed contact between every contig to bin
    and every bin, ~10^2 iterations of scalar LIL look-ups in the reference - through :class:`BinContactWeights` (the
    matrix is uploaded once per call).  The bipartite matching stays networkx's, the weights are the reference's bits, and
    every container is built by the same operations in the same order, because the reference's result depends on the
    iteration order of its sets and dicts (hence on PYTHONHASHSEED): under the same hash seed the bins are identical
    (tests/test_match_contigs.py).  No CPU path: without the CUDA library the weights raise.
    """
    import networkx as nx

    n_iterations = len(smg_iteration)
    bins, bin_of_contig, bin_markers = {}, {}, {}
    binned = []                                     # "binned_contigs_with_markers", utility.py:423
    n_bins = 0
    if n_iterations == 0:
        return bins, bin_of_contig, n_bins, bin_markers, binned

    operator = _operator
    for it in range(n_iterations):
        if it == 0:
            bins, bin_of_contig, bin_markers = _seed_bins(smg_iteration[0], contig_markers)
            binned.extend(smg_iteration[0])
            n_bins = len(bins)
            continue
        # contigs of this gene that no earlier iteration has binned (utility.py:447-449); a list made from a set
        # difference - the order of everything below follows from it
        already = set(binned).intersection(set(smg_iteration[it]))
        to_bin = list(set(smg_iteration[it]) - already)
        n_bins = len(bins)
        if not to_bin:
            continue
        # a bin is represented by its first contig (utility.py:453-456)
        representatives = []
        for b in range(n_bins):
            if bins[b][0] not in representatives:
                representatives.append(bins[b][0])
        candidates = []
        for contig in to_bin:
            if contig not in candidates:
                candidates.append(contig)

        # ---- utility.py:466-480 on the GPU: weight[r, b] = -(sum over the bin's members of E[contig_r, member]) / |bin|
        if operator is None:
            operator = BinContactWeights(normcc_matrix, device)
        bin_ptr, members = bins_to_arrays(bins, dict_contigRev)
        W = operator.weights([dict_contigRev[c] for c in to_bin], bin_ptr, members)

        graph = nx.Graph()
        graph.add_nodes_from(candidates, bipartite=0)
        graph.add_nodes_from(representatives, bipartite=1)
        weight_of = {}
        for r, contig in enumerate(to_bin):
            for b in range(n_bins):
                w = W[r, b]                         # numpy scalar of the matrix dtype, like the reference's running sum
                weight_of[(bins[b][0], contig)] = w
                graph.add_edge(bins[b][0], contig, weight=w)
        top = {node for node, attr in graph.nodes(data=True) if attr["bipartite"] == 0}
        bottom = set(graph) - top
        if not top:
            continue
        matching = nx.algorithms.bipartite.matching.minimum_weight_full_matching(graph, top, "weight")

        # ---- utility.py:503-544: a matched contig joins the bin when the contact is strong enough and it brings no
        # single-copy gene the bin already has; otherwise it is set aside
        set_aside = {}
        for rep in matching:
            if rep not in bin_of_contig:
                continue
            b = bin_of_contig[rep]
            contig = matching[rep]
            if contig in bins[b] or (rep, contig) not in weight_of:
                continue
            strong = -weight_of[(rep, contig)] >= w_intra
            if strong and not set(bin_markers[b]).intersection(set(contig_markers[contig])):
                bins[b].append(contig)
                bin_of_contig[contig] = b
                binned.append(contig)
                bin_markers[b] = list(set(bin_markers[b] + contig_markers[contig]))
            else:
                set_aside[contig] = (rep, b)
        # ---- utility.py:546-561: a contig set aside that has no contact above w_inter with ANY bin opens a new bin
        for contig in set_aside:
            lonely = True
            for rep in bottom:
                if -weight_of[(rep, contig)] > w_inter:
                    lonely = False
            if lonely:
                bins[n_bins] = [contig]
                bin_of_contig[contig] = n_bins
                bin_markers[n_bins] = contig_markers[contig]
                n_bins += 1
                binned.append(contig)
    return bins, bin_of_contig, n_bins, bin_markers, binned


def install_binning(*modules):
    """Rebind the name ``match_contigs`` that the reference's ``Script.pre_clustering`` / ``Script.final_cluster``
    imported (pre_clustering.py:8, final_cluster.py:19) to :func:`match_contigs_b200`."""
    for mod in modules:
        mod.match_contigs = match_contigs_b200


def precluster_b200(marker_contig_counts, marker_contigs, contig_markers, imputed_matrix, dict_contigRevLocal, intra, inter,
                    *, device=None, _operator=None):
    """Drop-in for ``PreCluster(marker_contig_counts, marker_contigs, contig_markers, imputed_matrix,
    dict_contigRevLocal, intra, inter)`` (/root/reference/Script/pre_clustering.py:10-63, called at ImputeCC.py:246, 342):
    the preliminary bins from the imputed matrix.  Returns ``(bins, bin_of_contig)``.

    The single-copy genes are taken in descending order of the number of contigs carrying them, ties by the total number
    of marker genes on those contigs (pre_clustering.py:11-44); an empty imputed matrix gives one bin per contig of the
    first gene (:47-53); otherwise :func:`match_contigs_b200` with the two percentile thresholds (:55-62), computed once
    and without the LIL / COO conversions (handoff.py)."""
    from .handoff import contact_thresholds
    order = {}
    for count in sorted(set(marker_contig_counts.values()), reverse=True):
        weight = {}
        for gene in marker_contig_counts:                       # dict order, as the reference walks it
            if marker_contig_counts[gene] == count:
                weight[gene] = sum(len(contig_markers[c]) for c in marker_contigs[gene])
        for gene, _ in sorted(weight.items(), key=lambda kv: kv[1], reverse=True):     # stable, like operator.itemgetter(1)
            order[len(order)] = marker_contigs[gene]
    thresholds = contact_thresholds(imputed_matrix, intra, inter)
    if thresholds is None:
        bins = {i: [c] for i, c in enumerate(order[0])}
        return bins, {c: i for i, c in enumerate(order[0])}
    bins, bin_of_contig, _, _, _ = match_contigs_b200(order, contig_markers, imputed_matrix, dict_contigRevLocal,
                                                      thresholds[0], thresholds[1], device=device, _operator=_operator)
    return bins, bin_of_contig
