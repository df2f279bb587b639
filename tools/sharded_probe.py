#!/usr/bin/env python
"""Developer probe (launch with torchrun): one sharded walk per rep, chunk by chunk, with CUDA-event times per chunk."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import imputecc_b200 as pkg
from imputecc_b200 import _lib, sharded, synth
from imputecc_b200.crwr import next_row_hints, poll_chunk

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "c3_sparse"
M, idx = synth.make_workload(synth.WORKLOADS[wl])
n, m = M.shape[0], idx.size
ranges = pkg.shard_rows(m, world, weights=np.diff(M.indptr)[idx])
w = pkg.Walker(n, m, dtype=np.float32, device=dev, transition_nnz_cap=M.nnz + n, row_range=ranges[rank], n_shards=world)
w.build_transition(M.indptr, M.indices, M.data, pkg.markers_first_order(n, idx))
peer = sharded.PeerExchange.get(dist, dist.group.WORLD, dev)
lib, h = w.lib, w.handle
for rep in range(3):
    if rank == 0 and os.environ.get("CRWR_TIMELINE"):
        lib.crwr_profile(h, 1)
    peer.attach(w)
    w.walk_init(0.5, 80)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    issued, hints, log, misses = 0, next_row_hints(0, 0, 80, 0), [], 0
    t_all = time.perf_counter()
    while True:
        chunk = poll_chunk(issued)
        w.set_row_hints(*hints)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        win = int(lib.crwr_step_window(h))
        e0.record()
        _lib.check(lib.crwr_walk_enqueue_peer(h, chunk, w._stream()))
        e1.record()
        issued += chunk
        st = w.poll()
        log.append((chunk, win, e0.elapsed_time(e1) * 1e3 / chunk))
        if int(st.capacity_overflow) & sharded.WINDOW_MISSED:
            misses += 1
            _lib.check(lib.crwr_walk_redo(h, w._stream()))
            issued = int(st.steps_done)
            continue
        if st.stop_reason != 0:
            break
        hints = w.row_hints(st, 80, issued)
    wall = (time.perf_counter() - t_all) * 1e3
    if rank == 0 and os.environ.get("CRWR_TIMELINE"):
        ms, nl = _lib.C.c_double(0), _lib.C.c_int64(0)
        lib.crwr_profile_read(h, _lib.C.byref(ms), _lib.C.byref(nl), w._stream())
    if rank == 0:
        print("rep %d: %d steps, wall %.3f ms, misses %d | chunks (steps, window form, us per step): %s" % (
            rep, st.steps_done, wall, misses, " ".join("%d/%d/%.0f" % c for c in log)), flush=True)
dist.barrier()
dist.destroy_process_group()
