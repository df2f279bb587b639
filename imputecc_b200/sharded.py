"""Sharded CRWR walk: restart rows split over the GPUs of one box.

Rows of the iterate never interact inside a step, so each rank walks a contiguous block of
restart rows against a replicated transition matrix.  The only coupled quantities are the
global percentile cut and ||dQ||_F; both travel as small uint64 buffers:

  exchange 0  all-reduce SUM   [8 counters incl. fixed-point delta^2 | level-0 histogram (4096)]
  exchange 1  all-reduce SUM   level-1 histogram (4096)
  exchange 2  all-gather       candidate block [count, successor key, <= 8192 keys]

after which every rank derives the same cut and the same stop decision on its own device - no
broadcast.  The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo
in the CPU tests) on views of the library's workspace.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import CapacityError, RetryError
from .crwr import (DEFAULT_MAX_STEPS, DEFAULT_TOL, WalkInfo, Walker, markers_first_order, shard_rows, stack_row_blocks,
                   symmetrise_device, validate_inputs, _require_cuda, _torch)


class CudaPhases:
    """The per-step phases of one shard, executed by libcrwr_b200 on this rank's GPU."""

    def __init__(self, walker: Walker):
        torch = _torch()
        self.w = walker
        lib, h = walker.lib, walker.handle
        self.views = []
        for which in range(3):
            ptr = lib.crwr_exchange_ptr(h, which)
            nbytes = int(lib.crwr_exchange_bytes(h, which))
            self.views.append(walker._ws_view(int(ptr), nbytes, torch.int64))

    def exchange(self, which):
        return self.views[which]

    def propagate(self):
        _lib.check(self.w.lib.crwr_step_propagate(self.w.handle, self.w._stream()))

    def hist(self, level):
        _lib.check(self.w.lib.crwr_step_hist(self.w.handle, level, self.w._stream()))

    def scan(self, level):
        _lib.check(self.w.lib.crwr_step_scan(self.w.handle, level, self.w._stream()))

    def gather(self):
        _lib.check(self.w.lib.crwr_step_gather(self.w.handle, self.w._stream()))

    def finish(self, gathered):
        ptr = C.c_void_p(gathered.data_ptr()) if gathered is not None else C.c_void_p(0)
        _lib.check(self.w.lib.crwr_step_finish(self.w.handle, ptr, self.w._stream()))

    def set_row_hints(self, hints):
        self.w.set_row_hints(*hints)

    def poll(self):
        return self.w.poll()

    def row_hints(self, st, perc, issued):
        return self.w.row_hints(st, perc, issued)


def _all_gather_blocks(dist, block, world, group):
    torch = _torch()
    out = torch.empty(world * block.numel(), dtype=block.dtype, device=block.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, block, group=group)
    else:
        parts = [out[r * block.numel():(r + 1) * block.numel()] for r in range(world)]
        dist.all_gather(parts, block, group=group)
    return out


class TorchComm:
    """The three collectives of a sharded step over a torch.distributed process group."""

    def __init__(self, dist, group):
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, group=self.group)

    def all_gather(self, block):
        return _all_gather_blocks(self.dist, block, self.world, self.group)

    def max_ints(self, values, like):
        torch = _torch()
        v = torch.tensor(list(values), dtype=torch.int64, device=like.device)
        self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX, group=self.group)
        return tuple(int(x) for x in v.cpu())


def run_sharded_walk(phases, comm, perc, max_steps):
    """Drive utility.py:281-304 across the shards behind ``comm``; every rank issues the same sequence.

    ``phases`` is a CudaPhases (the CPU tests pass an object with the same methods)."""
    from .crwr import next_row_hints, poll_chunk
    issued = 0
    hints = next_row_hints(0, 0, perc, 0)
    while True:
        chunk = min(poll_chunk(issued), 4)
        chunk = min(chunk, max_steps - issued)
        phases.set_row_hints(hints)
        for _ in range(chunk):
            phases.propagate()
            if issued >= 1:
                phases.hist(0)
            comm.all_reduce_sum(phases.exchange(0))
            if issued >= 1:
                phases.scan(0)
                phases.hist(1)
                comm.all_reduce_sum(phases.exchange(1))
                phases.scan(1)
                phases.gather()
                phases.finish(comm.all_gather(phases.exchange(2)))
            else:
                phases.finish(None)
            issued += 1
        st = phases.poll()
        if st.stop_reason != 0 or issued >= max_steps:
            return st
        # sizing hints affect speed only, never results; use the largest over the shards
        hints = comm.max_ints(phases.row_hints(st, perc, issued), phases.exchange(0))


class PeerExchange:
    """One symmetric-memory block per rank, mapped on every GPU of the group (NVLink peer access).

    Created once per (process group, device) and reused by successive walks; ``next_epoch_base`` hands every
    walk a fresh, strictly increasing range of barrier epochs so the flag words never need resetting."""

    _cache = []         # [(group object, device string, PeerExchange or None)]: keyed on the group itself, not id()

    def __init__(self, dist, group, device):
        torch = _torch()
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.dist, self.group = dist, group
        self.nbytes = int(lib.crwr_peer_block_bytes())
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.block = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.block.zero_()
        self.handle = symm_mem.rendezvous(self.block, group)
        offset = int(getattr(self.handle, "offset", 0))
        ptrs = [int(p) + offset for p in self.handle.buffer_ptrs]
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)      # device array of block pointers
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                                              # every block is zeroed before any use
        self.walks = 0

    @classmethod
    def get(cls, dist, group, device):
        """The exchange of (group, device), or None when it cannot be set up on EVERY rank (the ranks agree through
        an all-reduce, so either all of them use the peer path or all fall back to the collectives)."""
        torch = _torch()
        for g, d, ex in cls._cache:
            if g is group and d == str(device):
                return ex
        ex, ok = None, 1
        try:
            if not _peer_topology_ok(dist, group, device):
                ok = 0
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()):
            try:
                ex = cls(dist, group, device)
            except Exception:
                ex, ok = None, 0
            flag = torch.tensor([1 if ex is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if not int(flag.item()):
                ex = None
        cls._cache.append((group, str(device), ex))
        return ex

    def next_epoch_base(self) -> int:
        self.walks += 1
        return self.walks << 32

    def attach(self, walker):
        # no host barrier: the first device-side barrier of a walk waits up to two minutes for a late peer
        _lib.check(walker.lib.crwr_peer_attach(walker.handle, C.c_void_p(self.ptrs.data_ptr()), self.rank,
                                               C.c_uint64(self.next_epoch_base())))


def _peer_topology_ok(dist, group, device) -> bool:
    """All ranks of the group on this node, one GPU each, with peer access between every pair."""
    torch = _torch()
    world = dist.get_world_size(group)
    if world > torch.cuda.device_count():
        return False                      # more ranks than local GPUs: some live on another node
    import socket
    host = socket.gethostname()
    names = [None] * world
    dist.all_gather_object(names, (host, int(torch.device(device).index or 0)), group=group)
    if len({h for h, _ in names}) != 1 or len({d for _, d in names}) != world:
        return False
    me = int(torch.device(device).index or 0)
    return all(d == me or torch.cuda.can_device_access_peer(me, d) for _, d in names)


def enqueue_peer_step(walker, issued):
    """One step of a peer-attached shard on the walker's stream: the window form (one cross-GPU synchronisation) once
    the walk has settled, else the three-barrier form."""
    lib, h = walker.lib, walker.handle
    st = walker._stream()
    if lib.crwr_step_window(h):
        _lib.check(lib.crwr_step_propagate(h, st))
        _lib.check(lib.crwr_step_peer_tail(h, st))
        return
    _lib.check(lib.crwr_step_propagate(h, st))
    if issued >= 1:
        _lib.check(lib.crwr_step_peer_scan(h, 0, st))
        _lib.check(lib.crwr_step_hist(h, 1, st))
        _lib.check(lib.crwr_step_peer_scan(h, 1, st))
        _lib.check(lib.crwr_step_gather(h, st))
    else:
        _lib.check(lib.crwr_step_peer_scan(h, -1, st))
    _lib.check(lib.crwr_step_peer_finish(h, st))


WINDOW_MISSED = 4       # crwr_status.capacity_overflow bit: the select window of a sharded step missed


def run_sharded_walk_peer(walker, perc, max_steps):
    """The sharded walk with the exchanges fused into the consumer kernels (peer.cuh): no collective call in
    the step loop, only kernel launches on this rank's stream.  The walker must be peer-attached.

    Every rank sees the same window miss in the same step (the merged data are identical), so the ranks repeat that
    step in lockstep without talking to each other."""
    from .crwr import next_row_hints, poll_chunk
    lib, h = walker.lib, walker.handle
    issued = 0
    hints = next_row_hints(0, 0, perc, 0)
    while True:
        chunk = min(poll_chunk(issued), max_steps - issued)
        walker.set_row_hints(*hints)
        _lib.check(lib.crwr_walk_enqueue_peer(h, chunk, walker._stream()))
        issued += chunk
        status = walker.poll()
        if int(status.capacity_overflow) & WINDOW_MISSED:
            _lib.check(lib.crwr_walk_redo(h, walker._stream()))
            issued = int(status.steps_done)          # the steps queued behind the miss were skipped
            continue
        if status.stop_reason != 0 or issued >= max_steps:
            return status
        hints = walker.row_hints(status, perc, issued)      # local statistics: sizing affects speed only


def peer_exchange_available(dist, group) -> bool:
    """NVLink peer exchange needs NCCL ranks and torch symmetric memory (PeerExchange.get then checks that the ranks
    share a node with peer access and falls back, on all ranks together, when the rendezvous fails)."""
    if os.environ.get("CRWR_SHARDED_EXCHANGE", "peer") != "peer":
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception:
        return False
    return dist.get_backend(group) == "nccl" and 2 <= dist.get_world_size(group) <= 16


def gather_row_blocks(dist, group, ip, ix, dt, rows_of_rank, m):
    """All-gather per-rank CSR row blocks (device tensors) into one m x m CSR on every rank."""
    torch = _torch()
    world = dist.get_world_size(group)
    dev = dt.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[dist.get_rank(group)] = ix.numel()
    dist.all_reduce(counts, group=group)
    counts_h = [int(c) for c in counts.cpu()]
    pad = max(max(counts_h), 1)
    max_rows = max(r1 - r0 for r0, r1 in rows_of_rank)

    def padded(t, n, dtype):
        out = torch.zeros(n, dtype=dtype, device=dev)
        out[:t.numel()] = t
        return out

    g_ix = _all_gather_blocks(dist, padded(ix, pad, torch.int32), world, group)
    g_dt = _all_gather_blocks(dist, padded(dt, pad, dt.dtype), world, group)
    g_ip = _all_gather_blocks(dist, padded(ip, max_rows + 1, torch.int64), world, group)
    ips, ixs, dts = [], [], []
    base = 0
    for r, (r0, r1) in enumerate(rows_of_rank):
        rows = r1 - r0
        rip = g_ip[r * (max_rows + 1): r * (max_rows + 1) + rows + 1]
        ips.append(rip[:-1] + base)
        ixs.append(g_ix[r * pad: r * pad + counts_h[r]])
        dts.append(g_dt[r * pad: r * pad + counts_h[r]])
        base += counts_h[r]
    full_ip = torch.cat(ips + [torch.tensor([base], dtype=torch.int64, device=dev)])
    return full_ip, torch.cat(ixs), torch.cat(dts)


def crwr_impute_sharded(normcc, marker_idx_sorted, rp, perc, tol=DEFAULT_TOL, max_steps=DEFAULT_MAX_STEPS,
                        dtype=np.float32, device=None, return_info=False, group=None, iterate_cap=None):
    """crwr_impute with the restart rows sharded over the ranks of ``group`` (default: WORLD).

    Every rank passes the same inputs and receives the same imputed matrix."""
    torch = _torch()
    import torch.distributed as dist
    validate_inputs(normcc, marker_idx_sorted, rp, perc, check_values=False)
    if group is None:
        group = dist.group.WORLD
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = _require_cuda(device)
    M = sp.csr_matrix(normcc)
    if not M.has_canonical_format:
        M = M.copy()
        M.sum_duplicates()
    idx = np.asarray(marker_idx_sorted, dtype=np.int64)
    n, m = M.shape[0], idx.size
    if world > m:
        raise ValueError("more ranks than restart rows")
    new_of_old = markers_first_order(n, idx)
    ranges = shard_rows(m, world, weights=np.diff(M.indptr)[idx])
    r0, r1 = ranges[rank]
    w = Walker(n, m, dtype=dtype, device=device, transition_nnz_cap=int(M.nnz) + n, row_range=(r0, r1), n_shards=world,
               iterate_cap=iterate_cap)
    try:
        nnz_p = w.build_transition(M.indptr, M.indices, M.data, new_of_old)
        with torch.cuda.device(device):
            peer = PeerExchange.get(dist, group, device) if peer_exchange_available(dist, group) else None
            while True:
                # capacity overflows are all-reduced inside the step, so every rank sees the error in the same poll
                # and regrows in lockstep (the unsharded Walker.walk does the same)
                try:
                    if peer is not None:
                        peer.attach(w)
                    w.walk_init(rp, perc, tol, max_steps)
                    if peer is not None:
                        st = run_sharded_walk_peer(w, perc, max_steps)
                    else:
                        st = run_sharded_walk(CudaPhases(w), TorchComm(dist, group), perc, max_steps)
                    break
                except CapacityError as e:
                    if "candidate" in str(e) and w.regrows >= 1:
                        raise                 # more tied candidates than a shard block holds: a larger iterate buffer cannot help
                    w.cap *= 2
                    w.regrows += 1
                    w._recreate()
                except RetryError:
                    w.retries += 1
            ip, ix, dt = w.result_device()
            f_ip, f_ix, f_dt = gather_row_blocks(dist, group, ip, ix, dt, ranges, m)
            e_ip, e_ix, e_dt = symmetrise_device(f_ip, f_ix, f_dt, m)
        E = sp.csr_matrix((e_dt.cpu().numpy(), e_ix.cpu().numpy(), e_ip.cpu().numpy()), shape=(m, m))
        info = WalkInfo(int(st.steps_done), _lib.STOP_NAMES.get(int(st.stop_reason), "?"), w.trace() if return_info else [],
                        nnz_p, w.regrows)
    finally:
        w.close()
    return (E, info) if return_info else E
