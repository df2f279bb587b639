#!/usr/bin/env python
"""bench.py - CRWR imputation throughput on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3_sparse] [--dtype f32|f64]
    python bench.py --impl reference ...      # the reference's scipy algorithm on the host cores

One bench "step" = one complete CRWR imputation (normalise -> walk to its stop rule -> slice ->
symmetrise) of the named synthetic workload; the metric is BASELINE.json's "CRWR walk-steps/s":
walk steps executed per second of walk time.

  value  device-resident: transition matrix already in HBM, timed region = the walk itself
         (crwr_walk_init .. stop rule), CUDA events on the launching stream, max over ranks.
  e2e    the same metric through the public call crwr_impute() with HOST buffers: pinned
         NormCC CSR in, scipy matrix out, copies inside the timed region.
  roofline  the dominant kernel (propagate_group_kernel; propagate_kernel for step 0): algorithmic bytes per launch / its mean
         duration measured live with CUDA events, against MEASURED_PEAKS.json hbm_gbs.

N > 1 (torchrun): STRONG scaling of the same named workload (BASELINE config 3: 8,000 markers at 1/2/4/8
GPUs; `--workload c4_sparse --gpus 8` is config 4): the restart rows are sharded over the ranks, the
transition matrix is replicated, metric and unit stay "walk-steps/s" of the whole job.  Every line carries
`result_sha256` of the imputed matrix (canonical CSR: indptr | indices | data), so equal hashes at N = 1, 2, 4, 8
are the multi-GPU parity proof.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

def workload_for(name: str, world: int):
    """The named workload, whatever the number of GPUs (strong scaling)."""
    from imputecc_b200 import synth
    return synth.WORKLOADS[name]


def result_sha256(E) -> str:
    """sha256 of the canonical CSR form of the imputed matrix (int64 indptr | int32 indices | data)."""
    import hashlib
    import scipy.sparse as sp
    E = sp.csr_matrix(E).copy()
    E.sort_indices()
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(E.indptr.astype(np.int64)).tobytes())
    h.update(np.ascontiguousarray(E.indices.astype(np.int32)).tobytes())
    h.update(np.ascontiguousarray(E.data).tobytes())
    return h.hexdigest()


def walk_bytes(trace, nnz_p, n, s):
    """Algorithmic bytes of the walk (SURVEY.md §8d B_step) and the propagate kernel's share of it."""
    total = prop = 0
    for i, t in enumerate(trace):
        kept = t.get("nnz_kept", t["nnz_new"])
        w = nnz_p * (s + 4) + 4 * (n + 1)
        p = w + (s + 4) * (t["nnz_in"] + t["nnz_new"])          # stream W, read kept operand, write product
        total += p
        if i >= 1:
            total += (s + 4) * kept + s * t["nnz_new"]          # selection read + kept write
        prop += p
    return total, prop


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, dtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per propagate_kernel launch from the committed
    `ncu --set full` capture of this command (profiles/traffic.json), or None when not captured."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as fh:
            rec = json.load(fh).get("%s_%s" % (workload, dtype))
            return (rec["bytes_per_launch"], rec["source"]) if rec else (None, None)
    except Exception:
        return (None, None)


def pinned_csr(M):
    """scipy CSR whose arrays live in pinned host memory (asynchronous H2D)."""
    import scipy.sparse as sp
    import torch
    ip = torch.from_numpy(M.indptr.astype(np.int64)).pin_memory()
    ix = torch.from_numpy(M.indices.astype(np.int32)).pin_memory()
    dt = torch.from_numpy(M.data.astype(np.float64)).pin_memory()
    P = sp.csr_matrix((dt.numpy(), ix.numpy(), ip.numpy()), shape=M.shape, copy=False)
    P._pins = (ip, ix, dt)
    return P


# ---------------------------------------------------------------------------------------------
def _reference_walk():
    """The unmodified reference walk (Script.utility.random_walk_cpu) when the reference tree is importable
    (build container: /root/reference; a copy under baseline/_ref would do as well); None on the GPU box."""
    for root in ("/root/reference", os.path.join(REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(root, "Script")):
            try:
                if root not in sys.path:
                    sys.path.insert(0, root)
                import logging
                logging.disable(logging.CRITICAL)
                from Script.utility import random_walk_cpu
                return random_walk_cpu, root
            except Exception:
                continue
    return None, None


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on the host (scipy / numpy are single-threaded here, as in
    the reference): the unmodified `random_walk_cpu` when the reference tree is importable, else the oracle port
    (bit-pinned to it, tests/test_oracle.py).  Pre/post steps (imputation.py:112-119, :122-127) are timed beside."""
    if rank != 0:
        return
    from imputecc_b200 import synth
    from oracle import crwr_oracle as oracle
    wl = workload_for(args.workload, 1)
    M, idx = synth.make_workload(wl)
    np_dtype = np.float32               # the reference casts P to float32 (imputation.py:119)
    t0 = time.perf_counter()
    perm = oracle.marker_permutation(M.shape[0], idx)
    P = oracle.transition_matrix(M, perm, np_dtype)
    pre_s = time.perf_counter() - t0
    max_steps = args.ref_walk_steps
    ref_walk, ref_root = _reference_walk()
    kind = "reference" if (ref_walk is not None and max_steps >= 500) else "port"
    times, steps, post_s = [], 0, 0.0
    for i in range(args.warmup + args.steps):
        tr = oracle.WalkTrace()
        if kind == "reference":
            t0 = time.perf_counter()
            Q = ref_walk(P, 0.5, 80, 0.01, idx.size)
            dt = time.perf_counter() - t0
            if not steps or i == 0:
                oracle.crwr_walk(P, 0.5, 80, 0.01, idx.size, trace=tr)       # step count of the identical walk (untimed)
                n_steps = len(tr.steps)
        else:
            t0 = time.perf_counter()
            Q = oracle.crwr_walk(P, 0.5, 80, 0.01, idx.size, max_steps=max_steps, trace=tr)
            dt = time.perf_counter() - t0
            n_steps = len(tr.steps)
        if i >= args.warmup:
            times.append(dt)
            steps += n_steps
            t0 = time.perf_counter()
            E = oracle.symmetrise_normalise(Q)
            post_s = time.perf_counter() - t0
    total = sum(times)
    value = steps / total
    sample = "full walk" if max_steps >= 500 else "first %d walk steps of each walk" % max_steps
    line = {"impl": "reference", "metric": "crwr_walk_steps_per_s", "value": value, "unit": "walk-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, 1, "f32"),
            "cpu_baseline": {"value": value, "unit": "walk-steps/s", "cores": 1, "kind": kind,
                             "sample": "%s (%s); scipy.sparse + np.percentile are serial, 1 of %d host cores used" %
                                       (sample, "unmodified Script.utility.random_walk_cpu from %s" % ref_root if kind == "reference"
                                        else "oracle port, same scipy calls", os.cpu_count())},
            "pre_post_s": {"transition_matrix_s": pre_s, "symmetrise_normalise_s": post_s,
                           "note": "imputation.py:112-119 and :122-127 restated (oracle), outside the walk metric; the O(N^2) "
                                   "permutation-list loop of imputation.py:108-110 is not included"},
            "result_sha256": result_sha256(E),
            "e2e": {"value": value, "unit": "walk-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(wl, world, dtype):
    return {"workload": "%s: synthetic NormCC-like CSR, %d marker contigs on %d contigs, genome %d, intra-degree %d, "
                        "rwr-rp 0.5, rwr-thres 80" % (wl.name, wl.n_markers, wl.n_contigs, wl.genome_size, wl.intra_degree),
            "n_markers": wl.n_markers, "n_contigs": wl.n_contigs, "rp": 0.5, "perc": 80, "walk_dtype": dtype,
            "parallelism": "restart rows sharded over %d GPU(s), transition matrix replicated%s" % (
                world, "" if world == 1 else "; histogram/candidate exchange: %s" % os.environ.get("CRWR_SHARDED_EXCHANGE", "peer")),
            "l2": "L2 flushed (256 MiB write) between timed imputations"}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CRWR_BENCH_WORKLOAD", "c3_sparse"))
    ap.add_argument("--dtype", default=os.environ.get("CRWR_BENCH_DTYPE", "f32"), choices=["f32", "f64"])
    ap.add_argument("--ref-walk-steps", type=int, default=500, help="bound on walk steps per reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: imputecc_b200 has no CPU path (use --impl reference for the CPU arm)")
    import imputecc_b200 as pkg
    from imputecc_b200 import _lib, sharded, synth
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    np_dtype = np.float64 if args.dtype == "f64" else np.float32
    s = 8 if args.dtype == "f64" else 4

    wl = workload_for(args.workload, world)
    M, idx = synth.make_workload(wl)
    n, m = M.shape[0], idx.size
    new_of_old = pkg.markers_first_order(n, idx)
    ranges = pkg.shard_rows(m, world, weights=np.diff(M.indptr)[idx])
    walker = pkg.Walker(n, m, dtype=np_dtype, device=device, transition_nnz_cap=M.nnz + n,
                        row_range=ranges[rank] if world > 1 else None, n_shards=world)
    nnz_p = walker.build_transition(M.indptr, M.indices, M.data, new_of_old)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    comm = sharded.TorchComm(dist, dist.group.WORLD) if world > 1 else None

    use_peer = world > 1 and sharded.peer_exchange_available(dist, dist.group.WORLD)
    peer = sharded.PeerExchange.get(dist, dist.group.WORLD, device) if use_peer else None
    use_peer = peer is not None

    def one_walk():
        if world == 1:
            return walker.walk(0.5, 80)
        if use_peer:
            peer.attach(walker)                 # fresh barrier epochs for this walk
            walker.walk_init(0.5, 80)
            return sharded.run_sharded_walk_peer(walker, 80, 500)
        walker.walk_init(0.5, 80)
        return sharded.run_sharded_walk(sharded.CudaPhases(walker), comm, 80, 500)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    # ---- device-resident walk -------------------------------------------------------------
    for _ in range(args.warmup):
        one_walk()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.crwr_launch_count(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    walk_steps = 0
    for a, b in ev:
        flush.fill_(1)
        a.record()
        st = one_walk()
        b.record()
        walk_steps += int(st.steps_done)
    barrier()
    launches = int(lib.crwr_launch_count(0))
    walk_ms = sum(a.elapsed_time(b) for a, b in ev)
    trace = walker.trace()
    if rank == 0:
        print("trace:", [(t["nnz_in"], t["nnz_new"], t["max_row_len"], t["max_kept"], t["fallback_rows"]) for t in trace], file=sys.stderr)

    # ---- dominant kernel, live (CUDA events inside the library on the launching stream) ----
    def profiled_walk():
        lib.crwr_profile(walker.handle, 1)
        flush.fill_(1)
        one_walk()
        ms, nl = _lib.C.c_double(0), _lib.C.c_int64(0)
        _lib.check(lib.crwr_profile_read(walker.handle, _lib.C.byref(ms), _lib.C.byref(nl), walker._stream()))
        ph = (_lib.C.c_double * 3)()
        lib.crwr_profile_phases(walker.handle, ph)
        per_step = (_lib.C.c_double * 64)()
        n_steps = int(lib.crwr_profile_steps(walker.handle, per_step, 64))
        lib.crwr_profile(walker.handle, 0)
        return ms.value, int(nl.value), list(ph), [per_step[i] for i in range(n_steps)]

    steps_per_walk = len(trace)
    prop_ms, kern_launches, phases, step_ms_list = profiled_walk()
    kern_ms = prop_ms
    if world > 1:
        phases = [0.0, 0.0, 0.0]
    prop_ms_per_launch = prop_ms / max(steps_per_walk, 1)       # launches after the stop are no-ops (~0 ms)

    # ---- end to end through the public API, host buffers --------------------------------------
    Mp = pinned_csr(M)
    group = dist.group.WORLD if world > 1 else None
    e2e_times, e2e_steps, d2h = [], 0, 0
    for i in range(args.warmup + args.steps):
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        E, info = pkg.crwr_impute(Mp, idx, 0.5, 80, dtype=np_dtype, device=device, return_info=False, group=group), None
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_times.append(dt)
            e2e_steps += steps_per_walk
            d2h = E.data.nbytes + E.indices.nbytes + 8 * (m + 1)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- reduce over ranks (max time) -----------------------------------------------------------
    t = torch.tensor([walk_ms, sum(e2e_times) * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    walk_ms, e2e_ms = float(t[0]), float(t[1])

    value = walk_steps / (walk_ms * 1e-3)
    e2e_value = e2e_steps / (e2e_ms * 1e-3)
    total_bytes, prop_bytes = walk_bytes(trace, nnz_p, n, s)
    peak, peak_src = measured_hbm_peak()
    # rows are sharded: this rank's propagate launch moves ~1/world of the iterate plus all of W
    prop_bytes_launch = (prop_bytes / max(steps_per_walk, 1))
    if world > 1:
        w_bytes = nnz_p * (s + 4) + 4 * (n + 1)
        prop_bytes_launch = w_bytes + (prop_bytes_launch - w_bytes) / world
    prop_achieved = prop_bytes_launch / (prop_ms_per_launch * 1e-3) / 1e9 if prop_ms_per_launch > 0 else 0.0
    step_ms = walk_ms / max(walk_steps, 1)
    step_gbs = (total_bytes / max(steps_per_walk, 1)) / (step_ms * 1e-3) / 1e9
    if sum(t.get("fast_rows", 0) for t in trace) > 0:
        dom_name = ("cached-structure propagation: propagate_fast_big_kernel / propagate_fast_kernel + symbolic_kernel per walk step "
                    "(from the fourth step on; propagate_group_kernel before, propagate_kernel on step 0)")
    else:
        dom_name = "propagate_group_kernel (walk steps >= 1; step 0 runs propagate_kernel)"
    dom_bytes, dom_us, achieved = prop_bytes_launch, 1e3 * prop_ms_per_launch, prop_achieved
    # the settled walk alone (last third of the steps): the first steps run other kernels and, with cached structures, pay the build
    settled = step_ms_list[max(len(trace) - max(len(trace) // 3, 1), 0):len(trace)]
    settled_us = 1e3 * statistics.median(settled) if settled else 0.0
    settled_bytes = 0.0
    if settled and world == 1:
        t_last = trace[-1]
        settled_bytes = nnz_p * (s + 4) + 4 * (n + 1) + (s + 4) * (t_last["nnz_in"] + t_last["nnz_new"])
    h2d = M.indptr.size * 8 + M.indices.size * 4 + M.data.size * 8 + n * 4
    traffic, traffic_src = measured_traffic(args.workload, args.dtype) if world == 1 else (None, None)

    line = {
        "metric": "crwr_walk_steps_per_s", "value": value,
        "unit": "walk-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": walk_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": config_dict(wl, world, args.dtype),
        "walk_steps_per_imputation": steps_per_walk, "us_per_walk_step": 1e3 * step_ms,
        "imputation_wall_ms_e2e": e2e_ms / args.steps,
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "bytes_per_launch": dom_bytes, "us_per_launch": dom_us, "launches_per_imputation": kern_launches,
                     "propagate_kernel_alone": {"bytes_per_launch": prop_bytes_launch, "us_per_launch": 1e3 * prop_ms_per_launch,
                                                "achieved": prop_achieved, "frac": prop_achieved / peak,
                                                "note": "CUDA events around every launch, one profiled walk"},
                     "settled_step": {"bytes_per_launch": settled_bytes, "us_per_launch": settled_us,
                                      "achieved": settled_bytes / (settled_us * 1e-6) / 1e9 if settled_us > 0 else 0.0,
                                      "frac": (settled_bytes / (settled_us * 1e-6) / 1e9 / peak) if settled_us > 0 else 0.0,
                                      "note": "median of the propagation launches of the last third of the walk steps (B_prop of the last step)"},
                     "whole_step": {"bytes": total_bytes / max(steps_per_walk, 1), "us": 1e3 * step_ms,
                                    "achieved": step_gbs, "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0}},
        "e2e": {"value": e2e_value, "unit": "walk-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "step_phases_us": {"propagate+large_rows+scan": 1e3 * phases[0] / max(steps_per_walk - 1, 1),
                           "level1_histogram": 1e3 * phases[1] / max(steps_per_walk - 1, 1),
                           "gather+finish": 1e3 * phases[2] / max(steps_per_walk - 1, 1),
                           "note": "CUDA-event durations inside one profiled walk (unsharded walks only)"},
        "gpu_launches": launches,
        "result_sha256": result_sha256(E), "result_nnz": int(E.nnz),
        "clocks": sampler.summary(),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import crwr_oracle as oracle
        perm = oracle.marker_permutation(n, idx)
        P = oracle.transition_matrix(M, perm, np.float32)
        tr = oracle.WalkTrace()
        ref_walk, ref_root = _reference_walk()
        if ref_walk is not None:
            oracle.crwr_walk(P, 0.5, 80, 0.01, m, trace=tr)          # step count (untimed)
            t0 = time.perf_counter()
            ref_walk(P, 0.5, 80, 0.01, m)
            dt = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            oracle.crwr_walk(P, 0.5, 80, 0.01, m, trace=tr)
            dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": len(tr.steps) / dt, "unit": "walk-steps/s", "cores": 1,
                                "kind": "reference" if ref_walk is not None else "port",
                                "sample": "one full float32 walk (%d steps, %.2f s) of the same workload by %s; scipy.sparse and "
                                          "np.percentile are serial: 1 of %d host cores used" % (
                                              len(tr.steps), dt, "the unmodified Script.utility.random_walk_cpu" if ref_walk is not None
                                              else "the oracle port (same scipy calls; the Python reference does not travel to the GPU box)",
                                              os.cpu_count())}
    walker.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
