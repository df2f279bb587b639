// K3, settled-walk form: the percentile cut (utility.py:287) and the stop rules (utility.py:293-304) of one walk
// step as ONE single-CTA kernel behind the propagation kernels.
//
// The propagation epilogues (structure.cuh, propagate_group.cuh) have classified every value of the step against
// the window [lo, hi) that the previous step derived from its cut: `below` values under the window, the keys
// inside it as candidates, the smallest key at or above it as successor.  With n = number of stored values the
// percentile's order statistics x_(k), x_(k+1) lie among the candidates whenever below <= k < below + candidates
// (x_(k+1) may be the successor) - then the exact radix select over a few thousand keys replaces the two
// whole-array histogram passes, their scans and the gather pass of select.cuh.
//
// A window that misses (the cut moved further than predicted, or no window was set) is no error: the same CTA
// then runs the complete exact select over the step's value array by itself - slower (one SM streams the array
// from L2 up to three times), rare after the first few steps, and bit-identical by construction since both
// routes end in the same order statistics.
#pragma once
#include "select.cuh"
#include "structure.cuh"

namespace crwr {

struct TailArgs {
    FinishArgs fin;
    const void* vals;               // value array of this step's product (full select on a window miss)
    long long fixed_base;           // first entry of its fixed-slot region
    unsigned long long* cand;       // candidate block [count, successor, keys...]
    unsigned long long cand_cap;
};

// numpy percentile(), method 'linear': q and the virtual index live in the data dtype (see scan_level, level 0)
template <typename T>
__device__ __forceinline__ void virtual_index(WalkState* st, unsigned long long n) {
    const T q = div_rn((T)st->perc, (T)100);
    const T nm1 = (T)(n - 1);
    const T v = mul_rn(nm1, q);
    const T kf = floor(v);
    unsigned long long k = (unsigned long long)kf;
    int clamp = 0;
    T g = sub_rn(v, kf);
    if (v >= nm1 || k >= n - 1) { clamp = 1; k = n - 1; g = (T)0; }
    st->n_total = n;
    st->rank_k = k;
    st->gamma = (double)g;
    st->clamp_last = clamp;
}

// The complete exact select by one CTA (1024 threads): level-0 histogram (skipped when the epilogues filled it:
// no window was set this step), scan, level-1 histogram (skipped when pre-filled for the right bin), scan, gather.
template <typename T>
__device__ __noinline__ void full_select_one_cta(const TailArgs& a, unsigned char* scratch, bool hist0_valid) {
    typedef typename KeyOf<T>::type K;
    WalkState* st = a.fin.st;
    unsigned long long* H = a.fin.hist + kXExtras;
    unsigned int* sh = reinterpret_cast<unsigned int*>(scratch);
    const T* vals = reinterpret_cast<const T*>(a.vals);
    const ValRanges nvals = val_ranges(st, a.fixed_base);
    const int t = threadIdx.x;
    if (!hist0_valid) {
        for (int i = t; i < kSelectLevels * kSelectBins; i += blockDim.x) H[i] = 0ull;
        __syncthreads();
        hist_pass<T>(vals, nvals, (K)0, 0, sh, H);
        __syncthreads();
    }
    scan_level<T, 1024>(st, H, 0);
    const bool l1 = hist0_valid && st->pred_valid && st->prefix == st->pred_prefix0;
    __syncthreads();
    if (!l1) {
        for (int i = t; i < kSelectBins; i += blockDim.x) H[kSelectBins + i] = 0ull;
        __syncthreads();
        hist_pass<T>(vals, nvals, (K)st->prefix, 1, sh, H);
        __syncthreads();
    }
    scan_level<T, 1024>(st, H, 1);
    if (t == 0) { a.cand[0] = 0ull; a.cand[1] = ~0ull; }
    __syncthreads();
    gather_pass<T>(vals, nvals, (K)st->prefix, a.cand, a.cand_cap);
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(1024) tail_kernel(TailArgs a) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
    __shared__ int s_hit, s_free;
    pdl_wait();
    pdl_launch_dependents();
    WalkState* gst = a.fin.st;
    // the walk state in one coalesced read (thread 0 would otherwise make a dozen dependent L2 round trips)
    WalkState* st = finish_state_ptr(scratch);
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(gst);
        unsigned* dst = reinterpret_cast<unsigned*>(st);
        for (int i = threadIdx.x; i < (int)(sizeof(WalkState) / 4); i += blockDim.x) dst[i] = __ldcg(src + i);
    }
    __syncthreads();
    if (st->done) return;
    const bool win = st->win_on != 0;
    if (threadIdx.x == 0) {
        const unsigned long long n = st->nnz_exact;
        int hit = 0, free_bits = -1;
        if (win && n > 0) {
            virtual_index<T>(st, n);
            const unsigned long long c = a.cand[0], below = st->win_below, k = st->rank_k;
            if (c <= a.cand_cap && below <= k && k < below + c) {
                hit = 1;
                st->below = below;
                st->prefix = 0;
                st->bin_count = c;
                // keys inside [lo, hi) agree on every bit above the highest one in which lo and hi-1 differ
                const unsigned long long x = st->win_lo ^ (st->win_hi - 1ull);
                free_bits = x ? 64 - __clzll((long long)x) : 0;
                if (free_bits < 1) free_bits = 1;
            }
        }
        s_hit = hit;
        s_free = free_bits;
    }
    __syncthreads();
    if (s_hit) {
        // the window held: the histograms were not touched this step and stay clear
        finish_step_body<T>(a.fin, scratch, s_free, true, true);
        return;
    }
    // a miss: the complete select on the state in global memory, then the ordinary finish (which loads it again)
    if (threadIdx.x == 0) gst->win_miss_steps += 1;
    __syncthreads();
    full_select_one_cta<T>(a, scratch, !win);
    __syncthreads();
    finish_step_body<T>(a.fin, scratch, -1);
}

}  // namespace crwr
