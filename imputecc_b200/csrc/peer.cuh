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

namespace crwr {

constexpr int kPeerFlags = 16;                                       // max ranks
constexpr int kPeerX = kPeerFlags;                                    // [8 | 4096 | 4096]
constexpr int kPeerCand0 = kPeerX + kXExtras + 2 * kSelectBins;
constexpr int kPeerCandWords = kCandHeader + CRWR_SHARD_CAND;
constexpr int kPeerWords = kPeerCand0 + 2 * kPeerCandWords;

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
        const int off = kPeerX + kXExtras + level * kSelectBins;
        for (int b = t; b < kSelectBins; b += 1024) {
            unsigned long long s = 0;
            for (int r = 0; r < p.world; ++r) s += ld_peer(p.blocks[r] + off + b);
            local_xbuf[kXExtras + level * kSelectBins + b] = s;
        }
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
            const int off = kPeerX + kXExtras + kSelectBins;
            for (int b = t; b < kSelectBins; b += 1024) {
                unsigned long long s = 0;
                for (int r = 0; r < p.world; ++r) s += ld_peer(p.blocks[r] + off + b);
                local_xbuf[kXExtras + kSelectBins + b] = s;
            }
            __syncthreads();
            scan_level<T, 1024>(st, local_xbuf + kXExtras, 1);
            if (t == 0) st->hist1_ready = 1;
        } else {
            for (int b = t; b < kSelectBins; b += 1024) mine[kPeerX + kXExtras + kSelectBins + b] = 0ull;
            if (t == 0) st->hist1_ready = 0;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024) peer_finish_kernel(FinishArgs a, PeerInfo p, unsigned long long* gathered, int parity,
                                                           unsigned long long epoch) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
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

}  // namespace crwr
