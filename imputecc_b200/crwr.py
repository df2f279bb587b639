"""Host side of the B200 CRWR imputation: a drop-in for ImputeCC's imputation call.

Mirrors the reference interface (names, argument meaning, return types):

  reference                                                   here
  ----------------------------------------------------------  ---------------------------------
  Script/utility.py:270  random_walk_cpu(P, rp, perc, tol, m)  random_walk_b200(P, rp, perc, tol, m)
  Script/imputation.py:93 ImputeMatrix._imputation(self)       imputation_b200(self) / install()
  (the statements of _imputation as a free function)           crwr_impute(normcc, marker_idx, rp, perc)

PyTorch carries the buffers (device memory, pinned staging, streams, torch.distributed);
all arithmetic happens in libcrwr_b200.so (hand-written sm_100a kernels) behind the C ABI of
include/crwr_b200.h.  There is no CPU fallback: without the library or a CUDA device these
functions raise.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import CRWR_F32, CRWR_F64, CapacityError, Config, RetryError, Status, StepRecord

DEFAULT_TOL = 0.01          # imputation.py:120
DEFAULT_MAX_STEPS = 500     # utility.py:281
# walk steps enqueued between two status polls (steps after the stop rule fired are skipped on the device)
POLL_EVERY = int(__import__("os").environ.get("CRWR_POLL_EVERY", "16"))


def _torch():
    import torch
    return torch


def _require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("imputecc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("device must be a CUDA device")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return CRWR_F32
    if dt == np.float64:
        return CRWR_F64
    raise ValueError("walk dtype must be float32 or float64, got %r" % (dtype,))


# ---------------------------------------------------------------------------------------------
# host logic shared by the single- and multi-GPU paths (pure index work, no arithmetic)
# ---------------------------------------------------------------------------------------------

def markers_first_order(n_contigs: int, marker_idx_sorted) -> np.ndarray:
    """new_of_old[i]: position of contig i when markers come first (imputation.py:98-110).

    The reference builds the permutation list with an O(N^2) ``not in`` scan; this is the
    same order from a boolean mask.
    """
    marker_idx_sorted = np.asarray(marker_idx_sorted, dtype=np.int64)
    m = marker_idx_sorted.size
    rest = np.ones(n_contigs, dtype=bool)
    rest[marker_idx_sorted] = False
    # non-markers keep their order behind the m markers: position = m + (number of non-markers before i)
    new_of_old = np.cumsum(rest, dtype=np.int32)
    new_of_old += np.int32(m - 1)
    new_of_old[marker_idx_sorted] = np.arange(m, dtype=np.int32)
    return new_of_old


def validate_inputs(normcc, marker_idx_sorted, rp, perc, check_values=True):
    """Argument checks of the boundary (SURVEY.md §8b error conventions).

    ``check_values=False`` leaves the scan for negative / non-finite contacts to the device
    (crwr_build_transition reports it), saving a host pass over the matrix."""
    if not sp.issparse(normcc):
        raise ValueError("normcc must be a scipy sparse matrix")
    if normcc.shape[0] != normcc.shape[1]:
        raise ValueError("normcc must be square, got %r" % (normcc.shape,))
    idx = np.asarray(marker_idx_sorted)
    if idx.ndim != 1 or idx.size == 0:
        raise ValueError("at least one marker contig is required")
    if idx.min() < 0 or idx.max() >= normcc.shape[0]:
        raise ValueError("marker index out of range")
    if np.any(np.diff(idx) <= 0):
        raise ValueError("marker indices must be strictly increasing")
    if not (0.0 <= float(rp) <= 1.0):
        raise ValueError("rp must be in [0, 1]")
    if not (0.0 <= float(perc) <= 100.0):
        raise ValueError("Percentiles must be in the range [0, 100]")
    if check_values:
        data = normcc.data
        if data.size and (not np.all(np.isfinite(data)) or data.min() < 0):
            raise ValueError("contact values must be finite and non-negative")


def shard_rows(n_rows: int, n_shards: int, weights: Optional[np.ndarray] = None) -> List[tuple]:
    """Contiguous row ranges, balanced by ``weights`` (e.g. row degree) when given; never empty."""
    if n_shards < 1 or n_shards > n_rows:
        raise ValueError("need 1 <= n_shards <= n_rows")
    if weights is None:
        cuts = [(n_rows * r) // n_shards for r in range(n_shards + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64) + 1e-9
        cum = np.concatenate([[0.0], np.cumsum(w)])
        cuts = [0]
        for r in range(1, n_shards):
            c = int(np.searchsorted(cum, cum[-1] * r / n_shards))
            c = max(c, cuts[-1] + 1)
            c = min(c, n_rows - (n_shards - r))
            cuts.append(c)
        cuts.append(n_rows)
    return [(cuts[r], cuts[r + 1]) for r in range(n_shards)]


def default_iterate_cap(n_rows: int, n_contigs: int) -> int:
    return int(min(n_rows * n_contigs, max(1 << 22, 2048 * n_rows)))


def poll_chunk(issued: int) -> int:
    """Walk steps to enqueue before the next status poll.

    Steps 0 and 1 change the row shapes most, so the kernel sizing hints are refreshed after each;
    after step 3 the kept share has been observed and the walk runs in long chunks.  Steps enqueued
    after the stop rule fired are skipped on the device (a few empty launches each, ~7 us) - cheaper
    than an extra poll (~28 us: stream sync, status copy, relaunch gap) for walks of the usual 13-19 steps."""
    if issued < 2:
        return 1
    if issued < 4:
        return 2
    return POLL_EVERY


def next_row_hints(max_row_len: int, max_kept: int, perc: float, issued: int):
    """(longest output row, longest kept operand row) to expect in the next step.

    Sizing hints for the propagation kernel's per-row accumulator (crwr_set_row_hints).  Step 0
    reads Q = I; step 1 reads the un-cut step-0 rows and fans out two hops; from step 2 on the
    operand is the cut iterate, whose kept share is observed one step late.
    """
    if issued < 1:
        return 192, 32
    if issued < 2:
        # two hops: a row reaches up to (longest row) x (typical degree) columns; measured 6.5x (sparse) to 13x (dense
        # planted partitions) the step-0 row length.  A short hint sends rows to the slow large-row kernel.
        return 14 * max_row_len, max_row_len
    if issued < 3:
        # first cut steps: rows may still grow a little, the kept share is not yet observed, and the first structure
        # build consumes a superset of the kept entries (structure.cuh) that reaches more columns
        return 2 * max_row_len, int(max_row_len * min(1.0, 2.5 * (1.0 - perc / 100.0) + 0.05))
    return max_row_len, max_kept


def expanding_walk(prev, now) -> bool:
    """Host policy at the third poll: is the walk settling (the usual case at the reference's rwr-thres 80: from step 3 on
    the products grow by a few percent per step and the cut moves by less and less) or still expanding (rwr-thres 50 keeps
    half of every product: rows double from step to step until they are dense, and the cut falls five-fold per step)?

    ``prev`` / ``now``: (steps issued, entries of the last product, its cut) at the previous and at this poll.  For an
    expanding walk no select window can ever be set, and the single-CTA kernel of the settled-walk form would have to
    select over the whole value array by itself - 77 of 190 ms for 8,000 markers at rwr-thres 50 - so such a walk keeps
    the multi-CTA histogram select (crwr_hold_window_select).  (Measured and rejected for these walks: polling after every
    step with row hints scaled by the observed growth - the row-group kernel with tables for 20,000-column rows is slower
    than the large-row kernel the stale hints send them to.)"""
    steps = max(now[0] - prev[0], 1)
    if min(prev[1], now[1]) <= 0 or min(prev[2], now[2]) <= 0.0:
        return False
    growth = (float(now[1]) / float(prev[1])) ** (1.0 / steps)
    cut_ratio = (now[2] / prev[2]) ** (1.0 / steps)
    # steps 1 -> 3 of a settling walk: x1.2 per step (sparse) to x1.4 (dense planted partitions), cut x0.55 .. 0.66
    return growth > 1.5 and cut_ratio < 0.5


def stack_row_blocks(blocks: Sequence[sp.csr_matrix]) -> sp.csr_matrix:
    """Concatenate per-shard row blocks (same column count) into one CSR matrix."""
    return sp.vstack(list(blocks), format="csr")


@dataclass
class WalkInfo:
    steps: int
    stop_reason: str
    trace: list
    nnz_transition: int
    regrows: int = 0


# ---------------------------------------------------------------------------------------------
# the walker: owns one library handle and its workspace tensor
# ---------------------------------------------------------------------------------------------

class Walker:
    """One shard of a CRWR walk on one GPU (the whole walk when ``n_shards == 1``)."""

    def __init__(self, n_contigs: int, n_markers: int, dtype=np.float32, device=None,
                 transition_nnz_cap: Optional[int] = None, iterate_cap: Optional[int] = None,
                 row_range: Optional[tuple] = None, n_shards: int = 1):
        torch = _torch()
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.np_dtype = np.dtype(dtype)
        self.code = _dtype_code(dtype)
        self.torch_dtype = torch.float64 if self.code == CRWR_F64 else torch.float32
        self.n = int(n_contigs)
        self.m = int(n_markers)
        self.row_begin, self.row_end = (0, self.m) if row_range is None else (int(row_range[0]), int(row_range[1]))
        self.rows = self.row_end - self.row_begin
        self.n_shards = int(n_shards)
        self.pcap = int(transition_nnz_cap) if transition_nnz_cap else self.n * 32
        self.cap = int(iterate_cap) if iterate_cap else default_iterate_cap(self.rows, self.n)
        self.handle = C.c_void_p()
        self.ws = None
        self.regrows = 0
        self.retries = 0
        self._walk_args = None
        self.first_row_len = None      # longest restart row of the walk matrix, when known on the host (see walk())
        self.hint_override = None      # tests pin the kernel sizing through this
        self._create()

    # -- plumbing ---------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _cfg(self) -> Config:
        return Config(self.code, self.device.index, self.n, self.m, self.row_begin, self.row_end, self.pcap, self.cap,
                      self.n_shards, 0)

    def _create(self):
        torch = _torch()
        cfg = self._cfg()
        nbytes = self.lib.crwr_workspace_bytes(C.byref(cfg))
        if nbytes < 0:
            _lib.check(_lib.CRWR_ERR_INVALID)
        with torch.cuda.device(self.device):
            self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.crwr_create(C.byref(self.handle), C.byref(cfg), C.c_void_p(self.ws.data_ptr()), nbytes))

    def close(self):
        if self.handle:
            self.lib.crwr_destroy(self.handle)
            self.handle = C.c_void_p()
        self.ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _to_device(self, arr, dtype):
        torch = _torch()
        if isinstance(arr, torch.Tensor):
            t = arr.to(dtype) if arr.dtype != dtype else arr
        else:
            t = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)
        return t.to(self.device, non_blocking=True)

    def _ws_view(self, ptr: int, nbytes: int, dtype):
        off = ptr - self.ws.data_ptr()
        return self.ws[off:off + nbytes].view(dtype)

    # -- transition matrix ------------------------------------------------------------------
    def build_transition(self, indptr, indices, data, new_of_old) -> int:
        """imputation.py:112-119 on the device.  Arguments: CSR triplets of the NormCC matrix
        (numpy arrays or torch tensors, pinned for asynchronous upload) and markers_first_order()."""
        torch = _torch()
        self.first_row_len = None
        if isinstance(indptr, np.ndarray) and isinstance(new_of_old, np.ndarray) and self.n_shards == 1:
            # a row of P has the entries of the matrix row, minus a diagonal entry, plus a self-loop when isolated
            self.first_row_len = int(np.diff(indptr)[new_of_old < self.m].max()) + 1
        with torch.cuda.device(self.device):
            d_ip = self._to_device(indptr, torch.int64)
            d_ix = self._to_device(indices, torch.int32)
            d_dt = self._to_device(data, torch.float64)
            d_no = self._to_device(new_of_old, torch.int32)
            nnz = int(d_ix.numel())
            if nnz + self.n > self.pcap:
                self.pcap = nnz + self.n
                self._recreate()
            out = C.c_int64(0)
            _lib.check(self.lib.crwr_build_transition(self.handle, C.c_void_p(d_ip.data_ptr()), C.c_void_p(d_ix.data_ptr()),
                                                     C.c_void_p(d_dt.data_ptr()), nnz, C.c_void_p(d_no.data_ptr()),
                                                     C.byref(out), self._stream()))
        return int(out.value)

    def set_transition(self, P: sp.csr_matrix):
        """B-inner: P exactly as random_walk_cpu receives it (utility.py:270)."""
        torch = _torch()
        P = sp.csr_matrix(P)
        if P.shape != (self.n, self.n):
            raise ValueError("P must be %d x %d" % (self.n, self.n))
        if not P.has_canonical_format:
            P = P.copy()
            P.sum_duplicates()
        if P.nnz > self.pcap:
            self.pcap = int(P.nnz)
            self._recreate()
        self.first_row_len = int(np.diff(P.indptr[:self.m + 1]).max()) if self.n_shards == 1 and self.m > 0 else None
        with torch.cuda.device(self.device):
            d_ip = self._to_device(P.indptr.astype(np.int32, copy=False), torch.int32)
            d_ix = self._to_device(P.indices.astype(np.int32, copy=False), torch.int32)
            d_dt = self._to_device(P.data.astype(self.np_dtype, copy=False), self.torch_dtype)
            _lib.check(self.lib.crwr_set_transition(self.handle, C.c_void_p(d_ip.data_ptr()), C.c_void_p(d_ix.data_ptr()),
                                                   C.c_void_p(d_dt.data_ptr()), int(P.nnz), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def _transition_tensors(self):
        torch = _torch()
        nnz = int(self.lib.crwr_transition_nnz(self.handle))
        if nnz < 0:
            raise RuntimeError("no transition matrix")
        with torch.cuda.device(self.device):
            ip = torch.empty(self.n + 1, dtype=torch.int32, device=self.device)
            ix = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
            dt = torch.empty(max(nnz, 1), dtype=self.torch_dtype, device=self.device)
            _lib.check(self.lib.crwr_get_transition(self.handle, C.c_void_p(ip.data_ptr()), C.c_void_p(ix.data_ptr()),
                                                   C.c_void_p(dt.data_ptr()), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()
        return ip, ix[:nnz], dt[:nnz], nnz

    def transition(self) -> sp.csr_matrix:
        ip, ix, dt, _ = self._transition_tensors()
        return sp.csr_matrix((dt.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(self.n, self.n))

    def _recreate(self):
        """New workspace (capacity change); the transition matrix, if any, is carried over."""
        torch = _torch()
        have_p = self.handle and self.lib.crwr_transition_nnz(self.handle) >= 0
        saved = self._transition_tensors() if have_p else None
        self.close()
        self._create()
        if saved is not None:
            ip, ix, dt, nnz = saved
            _lib.check(self.lib.crwr_set_transition(self.handle, C.c_void_p(ip.data_ptr()), C.c_void_p(ix.data_ptr()),
                                                   C.c_void_p(dt.data_ptr()), nnz, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    # -- the walk -----------------------------------------------------------------------------
    def walk_init(self, rp, perc, tol=DEFAULT_TOL, max_steps=DEFAULT_MAX_STEPS):
        self._walk_args = (float(rp), float(perc), float(tol), int(max_steps))
        _lib.check(self.lib.crwr_walk_init(self.handle, float(rp), float(perc), float(tol), int(max_steps), self._stream()))

    def poll(self) -> Status:
        st = Status()
        _lib.check(self.lib.crwr_walk_poll(self.handle, C.byref(st), self._stream()))
        return st

    def walk(self, rp, perc, tol=DEFAULT_TOL, max_steps=DEFAULT_MAX_STEPS, first_row_len: Optional[int] = None) -> Status:
        """Run utility.py:281-304 to completion on this GPU (unsharded walk).

        ``first_row_len``: longest row of the walk matrix among the restart rows, when the caller knows it (step 0
        copies those rows, so the kernel sizing of step 1 needs no status poll after step 0)."""
        torch = _torch()
        if first_row_len is None:
            first_row_len = self.first_row_len
        with torch.cuda.device(self.device):
            while True:
                try:
                    return self._walk_once(rp, perc, tol, max_steps, first_row_len)
                except CapacityError:
                    if self.cap > 256 * default_iterate_cap(self.rows, self.n):    # not a sizing problem any more
                        raise
                    self.cap *= 2
                    self.regrows += 1
                    self._recreate()
                except RetryError:
                    self.retries += 1        # rows were deferred behind a step that queued no large-row kernel

    def _walk_once(self, rp, perc, tol, max_steps, first_row_len=None) -> Status:
        self.walk_init(rp, perc, tol, max_steps)
        lib, h = self.lib, self.handle
        issued = 0
        hints = next_row_hints(0, 0, perc, 0)
        if first_row_len is not None and first_row_len > 0:
            # steps 0 and 1 in one chunk: step 0 stores the restart rows of P (+ the restart entry), step 1 is sized from that
            hints = next_row_hints(int(first_row_len) + 1, 1, perc, 1)
        prev = None                 # what the poll before saw (expanding_walk)
        while True:
            # steps 0 and 1 change the row shapes most: poll after each; later run four at a time
            chunk = poll_chunk(issued)
            if issued == 0 and first_row_len is not None and first_row_len > 0:
                chunk = 2
            chunk = min(chunk, max_steps - issued)
            self.set_row_hints(*hints)
            _lib.check(lib.crwr_walk_enqueue(h, chunk, self._stream()))
            issued += chunk
            st = self.poll()
            if st.stop_reason != 0 or issued >= max_steps:
                return st
            hints = self.row_hints(st, perc, issued)
            now = (issued, int(st.nnz_iterate), float(st.last_cut))
            if issued == 4 and prev is not None and expanding_walk(prev, now):
                _lib.check(lib.crwr_hold_window_select(h, max_steps + 1))
            prev = now

    def set_row_hints(self, max_row_len: int, max_kept: int):
        if self.hint_override is not None:
            max_row_len, max_kept = self.hint_override
        _lib.check(self.lib.crwr_set_row_hints(self.handle, int(max_row_len), int(max_kept)))

    def row_hints(self, st: Status, perc: float, issued: int):
        return next_row_hints(int(st.max_row_len), int(st.max_kept), perc, issued)

    def trace(self) -> list:
        st = self.poll()
        n = max(int(st.steps_done), 0)
        buf = (StepRecord * max(n, 1))()
        _lib.check(self.lib.crwr_walk_trace(self.handle, buf, n, self._stream()))
        out = []
        for i in range(n):
            r = buf[i]
            out.append(dict(step=i, nnz_in=int(r.nnz_in), nnz_new=int(r.nnz_new), cut=float(r.cut),
                            cut_applied=bool(r.cut_applied), delta=float(r.delta), plateau_hits=int(r.plateau_hits),
                            fallback_rows=int(r.fallback_rows), max_row_len=int(r.max_row_len), max_kept=int(r.max_kept),
                            fast_rows=int(r.fast_rows)))
        for a, b in zip(out, out[1:]):
            a["nnz_kept"] = b["nnz_in"] if a["cut_applied"] else a["nnz_new"]
        return out

    # -- results ------------------------------------------------------------------------------
    def _rows_device(self, count_fn, write_fn):
        torch = _torch()
        with torch.cuda.device(self.device):
            nnz = C.c_int64(0)
            _lib.check(count_fn(self.handle, C.byref(nnz), self._stream()))
            n = int(nnz.value)
            ip = torch.empty(self.rows + 1, dtype=torch.int64, device=self.device)
            ix = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
            dt = torch.empty(max(n, 1), dtype=self.torch_dtype, device=self.device)
            _lib.check(write_fn(self.handle, C.c_void_p(ip.data_ptr()), C.c_void_p(ix.data_ptr()),
                                C.c_void_p(dt.data_ptr()), self._stream()))
        return ip, ix[:n], dt[:n]

    def result_device(self):
        """(indptr int64, indices int32, data) of this shard's rows of Q[:m,:m] (utility.py:306-307)."""
        return self._rows_device(self.lib.crwr_result_count, self.lib.crwr_result_write)

    def result(self) -> sp.csr_matrix:
        ip, ix, dt = self.result_device()
        return sp.csr_matrix((dt.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(self.rows, self.m))

    def iterate(self) -> sp.csr_matrix:
        """The full rows x N iterate after the last cut (tests, checkpoints)."""
        ip, ix, dt = self._rows_device(self.lib.crwr_iterate_count, self.lib.crwr_iterate_write)
        return sp.csr_matrix((dt.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(self.rows, self.n))


_PINNED = {}


def _pinned(dtype, n: int):
    """A pinned host staging buffer of at least n elements (cached per dtype: cudaHostAlloc is slow)."""
    torch = _torch()
    buf = _PINNED.get(dtype)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(int(n * 1.5), 1 << 16), dtype=dtype).pin_memory()
        _PINNED[dtype] = buf
    return buf


def download_csr(indptr, indices, data, shape) -> sp.csr_matrix:
    """Device CSR triplets -> scipy CSR with ONE stream synchronisation (three staged asynchronous copies)."""
    torch = _torch()
    if indptr.dtype == data.dtype or indices.dtype == data.dtype or not data.is_cuda:
        return sp.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=shape)
    views = []
    for t in (indptr, indices, data):
        host = _pinned(t.dtype, t.numel())[:t.numel()]
        host.copy_(t, non_blocking=True)
        views.append(host)
    torch.cuda.current_stream(data.device).synchronize()
    ip, ix, dt = (v.numpy().copy() for v in views)          # the staging buffers are reused by the next call
    return sp.csr_matrix((dt, ix, ip), shape=shape)


def symmetrise_device(indptr, indices, data, m: int):
    """imputation.py:122-127 on device tensors of an m x m CSR matrix with sorted rows."""
    torch = _torch()
    lib = _lib.load()
    device = data.device
    code = CRWR_F64 if data.dtype == torch.float64 else CRWR_F32
    nnz = int(indices.numel())
    with torch.cuda.device(device):
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        nbytes = int(lib.crwr_symmetrise_scratch_bytes(m, nnz))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
        # the library reads [0, nnz) only; give it valid pointers even when nnz == 0
        ix = indices if nnz else torch.zeros(1, dtype=torch.int32, device=device)
        dt = data if nnz else torch.zeros(1, dtype=data.dtype, device=device)
        out_nnz = C.c_int64(0)
        _lib.check(lib.crwr_symmetrise_count(code, m, C.c_void_p(indptr.data_ptr()), C.c_void_p(ix.data_ptr()),
                                             C.c_void_p(dt.data_ptr()), nnz, C.c_void_p(scratch.data_ptr()), nbytes,
                                             C.byref(out_nnz), stream))
        n = int(out_nnz.value)
        e_ip = torch.empty(m + 1, dtype=torch.int64, device=device)
        e_ix = torch.empty(max(n, 1), dtype=torch.int32, device=device)
        e_dt = torch.empty(max(n, 1), dtype=data.dtype, device=device)
        _lib.check(lib.crwr_symmetrise_write(code, m, C.c_void_p(indptr.data_ptr()), C.c_void_p(ix.data_ptr()),
                                             C.c_void_p(dt.data_ptr()), nnz, C.c_void_p(scratch.data_ptr()), nbytes,
                                             C.c_void_p(e_ip.data_ptr()), C.c_void_p(e_ix.data_ptr()),
                                             C.c_void_p(e_dt.data_ptr()), stream))
        # stream-ordered: scratch and inputs are returned to torch's allocator, which reuses them on this stream only
    return e_ip, e_ix[:n], e_dt[:n]


def symmetrise_normalise(Q, device=None) -> sp.csr_matrix:
    """E = Q + Q^T; D^-1/2 E D^-1/2 (imputation.py:122-127) for a scipy m x m matrix, on the GPU."""
    torch = _torch()
    device = _require_cuda(device)
    Q = sp.csr_matrix(Q)
    Q.sort_indices()
    m = Q.shape[0]
    ip = torch.from_numpy(Q.indptr.astype(np.int64)).to(device)
    ix = torch.from_numpy(Q.indices.astype(np.int32)).to(device)
    dt = torch.from_numpy(np.ascontiguousarray(Q.data)).to(device)
    e_ip, e_ix, e_dt = symmetrise_device(ip, ix, dt, m)
    with torch.cuda.device(device):
        return download_csr(e_ip, e_ix, e_dt, (m, m))


# ---------------------------------------------------------------------------------------------
# the reference-facing calls
# ---------------------------------------------------------------------------------------------

def random_walk_b200(P, rp, perc, tol, num_marker_contig, *, max_steps: int = DEFAULT_MAX_STEPS, device=None,
                     return_info: bool = False):
    """Drop-in for ``random_walk_cpu(P, rp, perc, tol, num_marker_contig)`` (utility.py:270).

    ``P``: scipy CSR N x N walk matrix, markers in rows/cols 0..m-1, float32 (reference) or
    float64.  Returns the m x m ``csc_matrix`` of the walk's dtype, like the reference.
    """
    if not sp.issparse(P) or P.shape[0] != P.shape[1]:
        raise ValueError("P must be a square scipy sparse matrix")
    m = int(num_marker_contig)
    if m < 1 or m > P.shape[0]:
        raise ValueError("num_marker_contig must be in [1, N]")
    if not (0.0 <= float(perc) <= 100.0):
        raise ValueError("Percentiles must be in the range [0, 100]")
    P = sp.csr_matrix(P)
    dtype = P.dtype if P.dtype in (np.float32, np.float64) else np.float64
    w = Walker(P.shape[0], m, dtype=dtype, device=device, transition_nnz_cap=max(int(P.nnz), 1))
    try:
        w.set_transition(P.astype(dtype, copy=False))
        st = w.walk(rp, perc, tol, max_steps)
        Q = w.result().tocsc()
        info = WalkInfo(int(st.steps_done), _lib.STOP_NAMES.get(int(st.stop_reason), "?"), w.trace() if return_info else [],
                        int(P.nnz), w.regrows)
    finally:
        w.close()
    return (Q, info) if return_info else Q


def crwr_impute(normcc, marker_idx_sorted, rp, perc, tol: float = DEFAULT_TOL, max_steps: int = DEFAULT_MAX_STEPS,
                dtype=np.float32, device=None, return_info: bool = False, group=None, iterate_cap: Optional[int] = None):
    """B-outer: NormCC matrix + sorted marker indices -> imputed m x m matrix.

    The statements of ``ImputeMatrix._imputation`` (imputation.py:112-127) on the GPU:
    transition matrix, walk, slice, symmetric normalisation.  Returns a scipy CSR matrix of the
    walk dtype (the caller applies ``.tolil()`` as imputation.py:129 does).  With a
    ``torch.distributed`` process group the restart rows are sharded over its ranks.  ``iterate_cap``
    overrides the initial capacity (entries) of the iterate buffers; they regrow when too small.
    """
    validate_inputs(normcc, marker_idx_sorted, rp, perc, check_values=False)
    if group is not None:
        from .sharded import crwr_impute_sharded
        return crwr_impute_sharded(normcc, marker_idx_sorted, rp, perc, tol, max_steps, dtype, device, return_info, group,
                                   iterate_cap=iterate_cap)
    M = normcc if isinstance(normcc, sp.csr_matrix) else sp.csr_matrix(normcc)
    if not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    idx = np.asarray(marker_idx_sorted, dtype=np.int64)
    n, m = M.shape[0], idx.size
    w = Walker(n, m, dtype=dtype, device=device, transition_nnz_cap=int(M.nnz) + n, iterate_cap=iterate_cap)
    try:
        torch = _torch()
        with torch.cuda.device(w.device):
            # the matrix goes up first (asynchronous from pinned memory); the markers-first order is computed on the
            # host while the copies run
            d_ip = w._to_device(M.indptr, torch.int64)
            d_ix = w._to_device(M.indices, torch.int32)
            d_dt = w._to_device(M.data, torch.float64)
            new_of_old = markers_first_order(n, idx)
            nnz_p = w.build_transition(d_ip, d_ix, d_dt, new_of_old)
            del d_ip, d_ix, d_dt
            # (a row of P has the entries of the matrix row, minus a diagonal entry, plus a self-loop when isolated)
            st = w.walk(rp, perc, tol, max_steps, first_row_len=int(np.diff(M.indptr)[idx].max()) + 1)
            ip, ix, dt = w.result_device()
            e_ip, e_ix, e_dt = symmetrise_device(ip, ix, dt, m)
            E = download_csr(e_ip, e_ix, e_dt, (m, m))
        info = WalkInfo(int(st.steps_done), _lib.STOP_NAMES.get(int(st.stop_reason), "?"), w.trace() if return_info else [],
                        nnz_p, w.regrows)
    finally:
        w.close()
    return (E, info) if return_info else E


def imputation_b200(self, dtype=np.float32, device=None, hand_off: str = "lil"):
    """Replacement for ``ImputeMatrix._imputation`` (imputation.py:93-129): same attributes read
    (normcc_matrix, contig_markers, dict_contigRev, contig_info, rwr_rp, rwr_perc), same tuple
    returned: (contig_local, dict_contigRevLocal, imputed lil_matrix).

    ``hand_off="csr"`` returns the imputed matrix as a CSR matrix instead: every consumer only calls ``.tocoo()``,
    ``.tolil()`` and ``.copy().tocsr()`` on it, and the LIL conversion alone costs ten times the imputation
    (imputecc_b200/handoff.py)."""
    if hand_off not in ("lil", "csr"):
        raise ValueError("hand_off must be 'lil' or 'csr'")
    names = list(self.contig_markers.keys())
    idx = np.sort(np.array([self.dict_contigRev[c] for c in names], dtype=np.int64))     # :101
    contig_local = self.contig_info[idx, :]                                             # :102
    rev_local = {name: i for i, name in enumerate(contig_local[:, 0])}                  # :104-106
    E = crwr_impute(self.normcc_matrix, idx, self.rwr_rp, self.rwr_perc, DEFAULT_TOL, DEFAULT_MAX_STEPS,
                    dtype=dtype, device=device)
    return contig_local, rev_local, (E.tolil() if hand_off == "lil" else E)            # :129


def install(imputation_module, mode: str = "outer", hand_off: str = "lil"):
    """Rebind the reference's entry point to the GPU path.

    ``imputation_module`` is the imported ``Script.imputation``.  mode "outer" replaces
    ``ImputeMatrix._imputation``; mode "inner" rebinds the name ``random_walk_cpu`` that
    ``Script.imputation`` imported (imputation.py:9), leaving pre/post-processing to scipy.
    """
    if hand_off not in ("lil", "csr"):
        raise ValueError("hand_off must be 'lil' or 'csr'")
    if mode == "outer":
        if hand_off == "lil":
            imputation_module.ImputeMatrix._imputation = imputation_b200
        else:
            imputation_module.ImputeMatrix._imputation = lambda self: imputation_b200(self, hand_off="csr")
    elif mode == "inner":
        imputation_module.random_walk_cpu = random_walk_b200
    else:
        raise ValueError("mode must be 'outer' or 'inner'")
