// The walk as ONE persistent cooperative kernel per chunk of steps (unsharded walks).
//
// A walk step is a chain of small dependent phases - propagate, (large rows), histogram scan,
// (level-1 histogram), candidate gather, finish - and at the sizes of this problem each phase lasts
// a few microseconds: launched as separate kernels, launch latency and last-CTA tails cost as much as
// the propagation itself.  Here the grid stays resident (one cooperative launch, CTAs co-resident on
// all SMs), phases are separated by grid barriers, the stop rules are evaluated on the device and the
// loop simply ends when they fire.  The phase bodies are the same device functions the stand-alone
// kernels (sharded walk, profiling) call.
#pragma once
#include <cooperative_groups.h>

#include "propagate.cuh"

namespace crwr {
namespace cg = cooperative_groups;

template <typename T>
struct WalkChunkArgs {
    PropArgs<T> prop;          // in/out/hist/fused_select are set per step
    IterBuf<T> buf[2];
    PropShape shape;
    DenseScratch<T> dense;
    long long dense_warps;
    FinishArgs fin;
    unsigned long long* hist;  // [2][4096]
    unsigned long long* cand_block;
    unsigned long long cand_cap;
    int n_steps;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PHASE_MARK(idx)                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                   \
        const unsigned long long now_ = global_ns();                             \
        st->phase_ns[idx] += now_ - t_prev;                                      \
        t_prev = now_;                                                           \
    }

__device__ __forceinline__ int ld_state(const int* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned ld_state(const unsigned* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) { return __ldcg(p); }

// Grid barrier + L1 invalidation on this SM (the iterate buffers and the walk state are rewritten by
// other SMs between phases and read with ordinary loads afterwards).
__device__ __forceinline__ void grid_barrier(cg::grid_group& grid) {
    grid.sync();
    if (threadIdx.x == 0) __threadfence();
    __syncthreads();
}

// Out-of-line copies of the phase bodies for the persistent kernel: inlined, the union of their register
// needs (255) would cap the residency of the propagate phase; the stand-alone kernels keep them inline.
template <typename T>
__device__ __noinline__ void dense_rows_ni(const PropArgs<T>& a, const DenseScratch<T>& sc, long long gw, unsigned int* win) {
    dense_rows<T>(a, sc, gw, win);
}
template <typename T>
__device__ __noinline__ void hist_pass_ni(const T* vals, long long n, typename KeyOf<T>::type prefix, int level, unsigned int* sh,
                                          unsigned long long* hist) {
    hist_pass<T>(vals, n, prefix, level, sh, hist);
}
template <typename T>
__device__ __noinline__ void gather_pass_ni(const T* vals, long long n, typename KeyOf<T>::type prefix, unsigned long long* block,
                                            unsigned long long cand_cap) {
    gather_pass<T>(vals, n, prefix, block, cand_cap);
}
template <typename T>
__device__ __noinline__ void finish_step_ni(const FinishArgs& a, unsigned char* scratch) {
    finish_step_body<T>(a, scratch);
}

template <typename T>
__device__ __forceinline__ void scan_tail_any(WalkState* st, unsigned long long* hist) {
    switch (blockDim.x) {
        case 32: scan_tail<T, 32>(st, hist); break;
        case 64: scan_tail<T, 64>(st, hist); break;
        case 128: scan_tail<T, 128>(st, hist); break;
        default: scan_tail<T, 256>(st, hist); break;
    }
}
template <typename T>
__device__ __forceinline__ void scan_level_any(WalkState* st, unsigned long long* hist, int level) {
    switch (blockDim.x) {
        case 32: scan_level<T, 32>(st, hist, level); break;
        case 64: scan_level<T, 64>(st, hist, level); break;
        case 128: scan_level<T, 128>(st, hist, level); break;
        default: scan_level<T, 256>(st, hist, level); break;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) walk_chunk_kernel(WalkChunkArgs<T> w) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    WalkState* st = w.prop.st;
    const int warps = blockDim.x >> 5;
    unsigned long long t_prev = global_ns();
    for (int s = 0; s < w.n_steps; ++s) {
        // the state is stable here: its last writer (the finish phase) precedes the barrier that ended the
        // previous step, so every CTA takes the same branch
        if (ld_state(&st->done)) break;
        const int step = ld_state(&st->step);
        const int parity = ld_state(&st->parity);
        const bool do_select = step >= 1;                      // no percentile cut on step 0 (utility.py:286)
        PropArgs<T> a = w.prop;
        a.in = w.buf[parity];
        a.out = w.buf[parity ^ 1];
        a.hist = do_select ? w.hist : nullptr;
        a.fused_select = 0;

        propagate_rows<T, true>(a, w.shape, smem_raw);
        PHASE_MARK(0)
        grid_barrier(grid);
        PHASE_MARK(1)

        if (ld_state(&st->fb_count) > 0u) {                     // rows that did not fit the shared-memory table
            unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw);
            for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) win[i] = 0u;
            __syncthreads();
            const long long gw = (long long)blockIdx.x * warps + (threadIdx.x >> 5);
            if (gw < w.dense_warps) dense_rows_ni<T>(a, w.dense, gw, win);
            __syncthreads();
            if (a.hist) hist_window_flush<T>(a.hist, win);
            grid_barrier(grid);
        }

        if (do_select) {
            if (blockIdx.x == 0) scan_tail_any<T>(st, w.hist);
            PHASE_MARK(2)
            grid_barrier(grid);
            PHASE_MARK(3)
            if (!ld_state(&st->hist1_ready)) {                  // the level-1 prediction missed: one more pass
                hist_pass_ni<T>(a.out.vals, (long long)ld_state(&st->cursor), (typename KeyOf<T>::type)ld_state(&st->prefix), 1,
                             reinterpret_cast<unsigned int*>(smem_raw), w.hist);
                grid_barrier(grid);
                if (blockIdx.x == 0) scan_level_any<T>(st, w.hist, 1);
                grid_barrier(grid);
            }
            gather_pass_ni<T>(a.out.vals, (long long)ld_state(&st->cursor), (typename KeyOf<T>::type)ld_state(&st->prefix),
                           w.cand_block, w.cand_cap);
            PHASE_MARK(4)
            grid_barrier(grid);
            PHASE_MARK(5)
        }
        if (blockIdx.x == 0) {
            FinishArgs f = w.fin;
            f.do_select = do_select ? 1 : 0;
            finish_step_ni<T>(f, smem_raw);
        }
        PHASE_MARK(6)
        grid_barrier(grid);
        PHASE_MARK(7)
    }
}

}  // namespace crwr
