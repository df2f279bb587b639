// Sharded walk over NVLink peer memory: the cross-GPU exchanges of a step are fused into the kernels
// that consume them, instead of three NCCL collectives.
//
// Every rank owns one symmetric block (torch symmetric memory; all blocks are mapped on every GPU):
//     [ flags[16] | X0 = 8 counters + level-0 histogram | X1 = level-1 histogram | CAND[2] ]   (uint64)
// A rank only ever WRITES its own block (histograms are filled there by its propagate / histogram
// kernels, candidates by its gather kernel) plus one flag word in each peer's block.  The consumer kernels
//   peer_scan   : cross-GPU flag barrier, then sum the peers' counters / histogram bins over NVLink into
//                 the local exchange buffer and locate the percentile bin (every rank computes the same sums)
//   peer_finish : barrier, copy the peers' candidate blocks, run the finish step, clear the own block
// replace all-reduce + scan and all-gather + finish.  Integer sums commute, so the cut and the stop
// decision are bit-identical on every rank and for every number of ranks.
//
// Hazards: a rank clears X0/X1 only in finish(t), after the barrier that every peer signals once it has
// read them (scan0/scan1 precede that signal); candidate blocks alternate by step parity, so the block
// peers read in finish(t) is next written by gather(t+2).
#pragma once
#include "select.cuh"
#include "tail.cuh"

namespace crwr {

constexpr int kPeerFlags = 16;                                       // max ranks
constexpr int kPeerX = kPeerFlags;                                    // [8 | 4096 | 4096]
constexpr int kPeerCand0 = kPeerX + kXExtras + 2 * kSelectBins;
constexpr int kPeerCandWords = kCandHeader + CRWR_SHARD_CAND;
// Window-select steps (peer_tail_kernel): every rank PUSHES its step counters and window candidates into the block of
// every peer - one region per (step parity, source rank) - and then raises one flag per peer; the consumer waits for the
// flags and reads local memory only.
constexpr int kPushCand = 8192;                                      // candidate keys a rank may push per step
constexpr int kPushBelow = kXExtras, kPushRaw = kXExtras + 1;       // header: 8 counters, values below the window, raw candidate count
constexpr int kPushBlock = 12;                                       // ... then a candidate block [count, successor, keys...]
constexpr int kPushWords = kPushBlock + kCandHeader + kPushCand;
constexpr int kPeerFlags2 = kPeerCand0 + 2 * kPeerCandWords;         // flags of the push exchange
constexpr int kPeerPush0 = kPeerFlags2 + kPeerFlags;
// Histogram steps (peer_scan_push_kernel): [8 counters | level-0 histogram | level-1 histogram] pushed the same way.
constexpr int kHPushWords = kXExtras + 2 * kSelectBins;
constexpr int kPeerFlags3 = kPeerPush0 + 2 * kPeerFlags * kPushWords;
constexpr int kPeerHPush0 = kPeerFlags3 + kPeerFlags;
constexpr int kPeerWords = kPeerHPush0 + 2 * kPeerFlags * kHPushWords;

struct PeerInfo {
    unsigned long long* const* blocks;   // device array [world]: every rank's symmetric block
    int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Sum over the ranks of word `idx` of every rank's block region `off`, for up to four words per thread at once: the
// remote loads of a batch (four ranks x four words) are all issued before the first of them is consumed - one
// dependent NVLink round trip per load was most of the exchange time.
__device__ __forceinline__ void peer_sum4(const PeerInfo& p, int off, const int idx[4], int n, unsigned long long out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0ull;
    for (int r0 = 0; r0 < p.world; r0 += 4) {
        unsigned long long v[4][4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const unsigned long long* base = p.blocks[r0 + rr < p.world ? r0 + rr : p.rank] + off;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[rr][k] = (r0 + rr < p.world && k < n) ? ld_peer(base + idx[k]) : 0ull;
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] += v[rr][k];
    }
}
// All kSelectBins words of region `off`, summed over the ranks, into dst (1024 threads, four words each).
__device__ __forceinline__ void peer_sum_bins(const PeerInfo& p, int off, unsigned long long* dst) {
    const int t = threadIdx.x;
    const int idx[4] = {t, t + 1024, t + 2048, t + 3072};
    unsigned long long s4[4];
    peer_sum4(p, off, idx, 4, s4);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[idx[k]] = s4[k];
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Cross-GPU barrier for one CTA per rank: publish `epoch` in every peer's flag slot for this rank, wait until
// every peer's epoch has arrived here.  Epochs only grow, so flags never need resetting.  The wait is
// bounded in wall time (a peer that died must not hang this GPU): 10 s inside a walk, two minutes at the first
// barrier of a walk (nothing on the host aligns the ranks before it: a peer may still be uploading its matrix).
// On time-out the walk is flagged, the caller skips its sums, and the finish step stops the walk.
__device__ __forceinline__ void peer_barrier(const PeerInfo& p, unsigned long long epoch, WalkState* st) {
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.world) {
        unsigned long long* remote = p.blocks[threadIdx.x] + p.rank;
        st_release_sys(remote, epoch);
        const unsigned long long* mine = p.blocks[p.rank] + threadIdx.x;
        const unsigned long long limit = st->step == 0 ? 120000000000ull : 10000000000ull;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(mine) < epoch) {
            if (global_ns() - t0 > limit) { st->peer_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
}

// level -1: counters only (step 0, no cut); 0: counters + level-0 bins + scan; 1: level-1 bins + scan.
template <typename T>
__global__ void __launch_bounds__(1024) peer_scan_kernel(WalkState* st, PeerInfo p, unsigned long long* local_xbuf, int level,
                                                         unsigned long long epoch) {
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    if (level == 1 && st->hist1_ready) return;      // level 1 was already resolved by the level-0 call (prediction hit)
    const int t = threadIdx.x;
    unsigned long long* mine = p.blocks[p.rank];
    if (level <= 0 && t == 0) {
        // publish this rank's step counters next to its level-0 histogram (what pack_exchange_kernel does)
        unsigned long long* extras = mine + kPeerX;
        extras[kXSq0] = st->sq_limb[0];
        extras[kXSq1] = st->sq_limb[1];
        extras[kXSq2] = st->sq_limb[2];
        extras[kXNnzIn] = st->nnz_in;
        extras[kXNnzNew] = st->nnz_exact;
        extras[kXCapOvf] = (unsigned long long)st->capacity_overflow;
        extras[kXSelOvf] = (unsigned long long)st->select_overflow;
        extras[kXFallback] = st->fb_count;
    }
    peer_barrier(p, epoch, st);
    if (st->peer_timeout) return;                   // unsynchronised data is never consumed; finish stops the walk
    if (level <= 0 && t < kXExtras) {
        unsigned long long s = 0;
        for (int r = 0; r < p.world; ++r) s += ld_peer(p.blocks[r] + kPeerX + t);
        local_xbuf[t] = s;
    }
    if (level < 0) return;
    {
        static_assert(kSelectBins == 4096, "peer_sum_bins covers 4096 bins with 1024 threads");
        peer_sum_bins(p, kPeerX + kXExtras + level * kSelectBins, local_xbuf + kXExtras + level * kSelectBins);
        __syncthreads();
        scan_level<T, 1024>(st, local_xbuf + kXExtras, level);
    }
    if (level == 0) {
        // every rank pre-filled its level-1 histogram for the level-0 bin predicted from the previous cut (the
        // prediction is part of the replicated state): on a hit level 1 is summed and scanned right here and the
        // level-1 pass with its barrier is skipped; on a miss the stale counts are dropped.
        const bool hit = st->pred_valid && st->prefix == st->pred_prefix0;
        __syncthreads();
        if (hit) {
            peer_sum_bins(p, kPeerX + kXExtras + kSelectBins, local_xbuf + kXExtras + kSelectBins);
            __syncthreads();
            scan_level<T, 1024>(st, local_xbuf + kXExtras, 1);
            if (t == 0) st->hist1_ready = 1;
        } else {
            for (int b = t; b < kSelectBins; b += 1024) mine[kPeerX + kXExtras + kSelectBins + b] = 0ull;
            if (t == 0) st->hist1_ready = 0;
        }
    }
}

// Level 0 of a histogram step, push form: every rank copies [8 counters | level-0 histogram | level-1 histogram
// pre-filled for the predicted bin] from its own block into its region of every peer's block (64 KB per peer, coalesced
// stores over NVLink), raises one flag per peer, waits for theirs and sums LOCAL memory - instead of pulling 2 x 4096
// bins from every peer with dependent remote loads behind two barriers.  When the level-0 bin predicted from the
// previous cut is the right one (it is, once the cut moves by a few percent), level 1 is resolved here as well and the
// level-1 pass with its exchange is skipped: two synchronisations per step (this one and the finish) instead of three.
template <typename T>
__global__ void __launch_bounds__(1024) peer_scan_push_kernel(WalkState* st, PeerInfo p, unsigned long long* local_xbuf, int parity,
                                                              unsigned long long epoch) {
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    const int t = threadIdx.x;
    unsigned long long* mine = p.blocks[p.rank];
    if (t == 0) {
        unsigned long long* extras = mine + kPeerX;
        extras[kXSq0] = st->sq_limb[0];
        extras[kXSq1] = st->sq_limb[1];
        extras[kXSq2] = st->sq_limb[2];
        extras[kXNnzIn] = st->nnz_in;
        extras[kXNnzNew] = st->nnz_exact;
        extras[kXCapOvf] = (unsigned long long)st->capacity_overflow;
        extras[kXSelOvf] = (unsigned long long)st->select_overflow;
        extras[kXFallback] = st->fb_count;
    }
    __syncthreads();
    // ---- push
    const int reg = kPeerHPush0 + (parity * kPeerFlags + p.rank) * kHPushWords;
    for (int i0 = 0; i0 < kHPushWords; i0 += 4 * 1024) {
        unsigned long long v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = i0 + k * 1024 + t; v[k] = i < kHPushWords ? mine[kPeerX + i] : 0ull; }
        for (int r = 0; r < p.world; ++r) {
            unsigned long long* dst = p.blocks[r] + reg;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int i = i0 + k * 1024 + t; if (i < kHPushWords) dst[i] = v[k]; }
        }
    }
    __syncthreads();
    if (t == 0) __threadfence_system();
    __syncthreads();
    if (t < p.world) {
        st_release_sys(p.blocks[t] + kPeerFlags3 + p.rank, epoch);
        const unsigned long long* flag = mine + kPeerFlags3 + t;
        const unsigned long long limit = st->step <= 1 ? 120000000000ull : 10000000000ull;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < epoch) {
            if (global_ns() - t0 > limit) { st->peer_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (t == 0) __threadfence_system();
    __syncthreads();
    if (st->peer_timeout) return;
    // ---- sum local memory
    const unsigned long long* regs = mine + kPeerHPush0 + (long long)parity * kPeerFlags * kHPushWords;
    for (int i = t; i < kXExtras + kSelectBins; i += 1024) {
        unsigned long long sum = 0;
        for (int r = 0; r < p.world; ++r) sum += regs[(long long)r * kHPushWords + i];
        local_xbuf[i] = sum;
    }
    __syncthreads();
    scan_level<T, 1024>(st, local_xbuf + kXExtras, 0);
    const bool hit = st->pred_valid && st->prefix == st->pred_prefix0;
    __syncthreads();
    if (hit) {
        for (int i = kXExtras + kSelectBins + t; i < kHPushWords; i += 1024) {
            unsigned long long sum = 0;
            for (int r = 0; r < p.world; ++r) sum += regs[(long long)r * kHPushWords + i];
            local_xbuf[i] = sum;
        }
        __syncthreads();
        scan_level<T, 1024>(st, local_xbuf + kXExtras, 1);
        if (t == 0) st->hist1_ready = 1;
    } else {
        for (int b = t; b < kSelectBins; b += 1024) mine[kPeerX + kXExtras + kSelectBins + b] = 0ull;
        if (t == 0) st->hist1_ready = 0;
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) peer_finish_kernel(FinishArgs a, PeerInfo p, unsigned long long* gathered, int parity,
                                                           unsigned long long epoch) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
    pdl_wait();
    pdl_launch_dependents();
    if (a.st->done) return;
    unsigned long long* mine = p.blocks[p.rank];
    // always a barrier: on step 0 it protects the counters a slower peer may still be summing
    peer_barrier(p, epoch, a.st);
    if (a.do_select && !a.st->peer_timeout) {
        const int cand_off = kPeerCand0 + parity * kPeerCandWords;
        for (int r = 0; r < p.world; ++r) {
            const unsigned long long* src = p.blocks[r] + cand_off;
            unsigned long long* dst = gathered + (long long)r * kPeerCandWords;
            const unsigned long long cnt = ld_peer(src);
            const unsigned long long n = cnt < (unsigned long long)CRWR_SHARD_CAND ? cnt : (unsigned long long)CRWR_SHARD_CAND;
            for (unsigned long long i = threadIdx.x; i < kCandHeader + n; i += blockDim.x) dst[i] = ld_peer(src + i);
        }
        __syncthreads();
        a.blocks = gathered;
        a.n_blocks = p.world;
        a.block_stride = kPeerCandWords;
    }
    // the candidate block of the NEXT step is reset here (peers finished reading it two barriers ago)
    a.own_block = mine + kPeerCand0 + (parity ^ 1) * kPeerCandWords;
    finish_step_body<T>(a, scratch);
    // every peer has read this rank's counters and histograms of the step (they signalled the barrier above,
    // or the scan barrier on step 0, only after doing so): clear them for the next step
    for (int i = threadIdx.x; i < kXExtras + 2 * kSelectBins; i += blockDim.x) mine[kPeerX + i] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// Settled sharded walk: the percentile cut of a step from ONE cross-GPU synchronisation.  The propagation epilogues
// of every rank have classified their values against the window [lo, hi) that the previous step derived from its cut
// (replicated state, identical on all ranks): values below it counted, keys inside it collected, smallest key above
// it kept.  This kernel (one CTA per rank)
//   1. pushes [8 step counters | below | candidate count | successor | keys] into its region of every peer's block
//      (stores over NVLink, coalesced), fences, and raises its flag in every peer's block;
//   2. waits for the flags of all peers - from here on everything it reads is local memory;
//   3. sums the counters, concatenates the candidates and - when the order statistics of the percentile lie among
//      them - selects the cut and runs the finish step, exactly like the unsharded tail (tail.cuh).
// Integer sums and the union of the candidate sets do not depend on the rank, so every rank derives the same cut and
// the same stop decision.  A window that misses (or was not set, or overflowed a region) is no error: every rank sees
// the same miss, rewinds its step counters and stops with `win_missed`; the host then queues this step again the
// three-barrier way (peer_scan / peer_finish above).
// Regions alternate by step parity: a rank can only push step t+2 after every peer has raised the flag of step t+1,
// which a peer does after it has finished reading step t.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) peer_tail_kernel(FinishArgs a, PeerInfo p, unsigned long long* local_xbuf,
                                                         unsigned long long* my_cand, unsigned long long* next_cand, int parity,
                                                         unsigned long long epoch) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
    __shared__ int s_hit, s_free;
    const int t = threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    WalkState* gst = a.st;
    WalkState* st = finish_state_ptr(scratch);
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(gst);
        unsigned* dst = reinterpret_cast<unsigned*>(st);
        for (int i = t; i < (int)(sizeof(WalkState) / 4); i += blockDim.x) dst[i] = __ldcg(src + i);
    }
    __syncthreads();
    if (st->done) return;
    unsigned long long* mine = p.blocks[p.rank];
    // ---- 1. push
    const unsigned long long raw = my_cand[0];
    const unsigned long long npush = raw < (unsigned long long)kPushCand ? raw : (unsigned long long)kPushCand;
    const int reg = kPeerPush0 + (parity * kPeerFlags + p.rank) * kPushWords;
    unsigned long long hdr = 0ull;
    switch (t) {
        case kXSq0: hdr = st->sq_limb[0]; break;
        case kXSq1: hdr = st->sq_limb[1]; break;
        case kXSq2: hdr = st->sq_limb[2]; break;
        case kXNnzIn: hdr = st->nnz_in; break;
        case kXNnzNew: hdr = st->nnz_exact; break;
        case kXCapOvf: hdr = (unsigned long long)st->capacity_overflow; break;
        case kXSelOvf: hdr = (unsigned long long)st->select_overflow; break;
        case kXFallback: hdr = st->fb_count; break;
        case kPushBelow: hdr = st->win_below; break;
        case kPushRaw: hdr = raw; break;
        case kPushBlock: hdr = npush; break;
        case kPushBlock + 1: hdr = my_cand[1]; break;
        default: break;
    }
    for (int r = 0; r < p.world; ++r) {
        unsigned long long* dst = p.blocks[r] + reg;
        if (t < kPushBlock + kCandHeader) dst[t] = hdr;
        for (unsigned long long i = t; i < npush; i += blockDim.x) dst[kPushBlock + kCandHeader + i] = my_cand[kCandHeader + i];
    }
    __syncthreads();
    if (t == 0) __threadfence_system();
    __syncthreads();
    // ---- 2. one flag per peer, then wait for theirs
    if (t < p.world) {
        st_release_sys(p.blocks[t] + kPeerFlags2 + p.rank, epoch);
        const unsigned long long* flag = mine + kPeerFlags2 + t;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) < epoch) {
            if (global_ns() - t0 > 10000000000ull) { st->peer_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (t == 0) __threadfence_system();
    __syncthreads();
    // ---- 3. merge (local memory), decide
    const unsigned long long* regs = mine + kPeerPush0 + (long long)parity * kPeerFlags * kPushWords;
    if (t < kXExtras) {
        unsigned long long sum = 0;
        for (int r = 0; r < p.world; ++r) sum += regs[(long long)r * kPushWords + t];
        local_xbuf[t] = sum;
    }
    if (t == 0) {
        unsigned long long below = 0, c = 0, n = 0;
        int ovf = 0;
        for (int r = 0; r < p.world; ++r) {
            const unsigned long long* rg = regs + (long long)r * kPushWords;
            below += rg[kPushBelow];
            c += rg[kPushBlock];
            n += rg[kXNnzNew];
            if (rg[kPushRaw] > (unsigned long long)kPushCand) ovf = 1;
        }
        int hit = 0, free_bits = -1;
        if (st->win_on && n > 0 && !ovf && !st->peer_timeout) {
            virtual_index<T>(st, n);
            const unsigned long long k = st->rank_k;
            if (below <= k && k < below + c) {
                hit = 1;
                st->below = below;
                st->prefix = 0;
                st->bin_count = c;
                const unsigned long long x = st->win_lo ^ (st->win_hi - 1ull);
                free_bits = x ? 64 - __clzll((long long)x) : 0;
                if (free_bits < 1) free_bits = 1;
            }
        }
        if (st->peer_timeout) hit = 1;          // the finish step stops the walk with the error
        s_hit = hit;
        s_free = free_bits;
    }
    __syncthreads();
    if (s_hit) {
        a.xch = local_xbuf;
        a.blocks = const_cast<unsigned long long*>(regs) + kPushBlock;
        a.n_blocks = p.world;
        a.block_stride = kPushWords;
        a.cand_cap = (unsigned long long)kPushCand;
        a.own_block = next_cand;                // the candidate block of the next step (its last readers are long gone)
        finish_step_body<T>(a, scratch, s_free, true, true);
        if (t == 0) { my_cand[0] = 0ull; my_cand[1] = ~0ull; }
        return;
    }
    // ---- a miss: rewind the step (every rank does), stop with win_missed; the host queues the step again
    if (t == 0) {
        st->win_miss_steps += 1;
        st->cursor = 0; st->nnz_exact = 0; st->nnz_in = 0; st->max_row_len = 0; st->max_kept = 0;
        st->row_counter = 0; st->fb_counter = 0; st->fb_count = 0;
        st->sq_limb[0] = 0; st->sq_limb[1] = 0; st->sq_limb[2] = 0;
        st->slow_count = 0; st->slow2_count = 0; st->fast_rows = 0; st->built_rows = 0; st->win_below = 0;
        st->tick_dense = 0; st->tick_hist1 = 0; st->tick_gather = 0; st->tick_prop = 0;
        st->scan_done = 0; st->hist1_ready = 0; st->unordered_rows = 0;
        st->win_on = 0;
        st->win_missed = 1;
        st->done = 1;
        st->stop_reason = CRWR_STOP_RUNNING;
        my_cand[0] = 0ull; my_cand[1] = ~0ull;
    }
    __syncthreads();
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(st);
        unsigned* dst = reinterpret_cast<unsigned*>(gst);
        for (int i = t; i < (int)(sizeof(WalkState) / 4); i += blockDim.x) dst[i] = src[i];
    }
    // the histograms may hold this attempt's counts (no window was set): clear them for the second attempt
    for (int i = t; i < kXExtras + 2 * kSelectBins; i += blockDim.x) mine[kPeerX + i] = 0ull;
}

// After a polled miss: the walk goes on (the host queues the missed step again, histogram select).
__global__ void peer_redo_kernel(WalkState* st) {
    if (threadIdx.x == 0 && st->win_missed) { st->win_missed = 0; st->done = 0; }
}

}  // namespace crwr
