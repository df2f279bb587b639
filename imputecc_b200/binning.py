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
