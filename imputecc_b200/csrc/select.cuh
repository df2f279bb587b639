// K3: the per-step percentile cut  np.percentile(Q.data, perc)  (utility.py:287) as an exact
// radix select on order-preserving integer keys, plus the step's stop logic (utility.py:293-304).
//
//   level 0 histogram (top 12 key bits, all values)  -> scan: total n, virtual index, target bin
//   level 1 histogram (next 12 bits, values in bin)  -> scan: 24-bit prefix of x_(k)
//   gather: values whose 24-bit prefix matches -> candidate list; min key above the bin -> successor
//   finish: rank candidates -> x_(k), x_(k+1) -> numpy 'linear' interpolation -> cut; delta; stop rules
//
// Histograms are plain integer counts, so a sharded walk only has to SUM them across ranks
// (and concatenate candidates) for every rank to derive the identical cut.
#pragma once
#include "common.cuh"

namespace crwr {

// True in exactly one CTA of the grid: the last one to arrive.  All threads of every CTA call it.
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Block-wide (BLOCK threads): locate the bin of `hist[level]` that holds rank_k; level 0 first derives
// n, the virtual index and gamma from the total count.  hist holds the GLOBAL counts of the level.
template <typename T, int BLOCK>
__device__ __noinline__ void scan_level(WalkState* st, const unsigned long long* hist, int level) {
    constexpr int BPT = kSelectBins / BLOCK;
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long s_total;
    const unsigned long long* h = hist + (size_t)level * kSelectBins;
    const int t = threadIdx.x;
    // a thread's bins stay in registers when there are few of them; small CTAs (many bins per thread)
    // re-read them instead of holding a 128-entry register array
    constexpr bool kInRegs = BPT <= 32;
    unsigned long long c[kInRegs ? BPT : 1];
    unsigned long long mine = 0;
    if constexpr (kInRegs) {
#pragma unroll
        for (int i = 0; i < BPT; ++i) { c[i] = __ldcg(h + t * BPT + i); mine += c[i]; }
    } else {
#pragma unroll 8
        for (int i = 0; i < BPT; ++i) mine += __ldcg(h + t * BPT + i);
    }
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long v = __shfl_up_sync(kFull, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        unsigned long long w = t < BLOCK / 32 ? warp_tot[t] : 0ull;
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long v = __shfl_up_sync(kFull, wi, o);
            if (t >= o) wi += v;
        }
        if (t < BLOCK / 32) warp_tot[t] = wi - w;   // exclusive
        if (t == 31) s_total = wi;
    }
    __syncthreads();
    const unsigned long long excl = warp_tot[t >> 5] + incl - mine;

    if (level == 0) {
        if (t == 0) {
            // numpy percentile(), method 'linear': q and the virtual index live in the data dtype
            const unsigned long long n = s_total;
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
            st->below = 0;
            st->prefix = 0;
        }
        __syncthreads();
    }
    const unsigned long long below0 = st->below;
    const unsigned long long prefix0 = st->prefix;
    const unsigned long long target = st->rank_k - below0;      // rank inside the current bin set
    __syncthreads();                                            // all reads precede the single write
    if (target >= excl && target < excl + mine) {          // exactly one thread: the bin lies in its range
        unsigned long long run = excl;
        if constexpr (kInRegs) {
#pragma unroll
            for (int i = 0; i < BPT; ++i) {
                if (target >= run && target < run + c[i]) {
                    st->below = below0 + run;
                    st->prefix = (prefix0 << kSelectBits) | (unsigned long long)(t * BPT + i);
                    st->bin_count = c[i];
                }
                run += c[i];
            }
        } else {
            for (int i = 0; i < BPT; ++i) {
                const unsigned long long ci = __ldcg(h + t * BPT + i);
                if (target < run + ci) {
                    st->below = below0 + run;
                    st->prefix = (prefix0 << kSelectBits) | (unsigned long long)(t * BPT + i);
                    st->bin_count = ci;
                    break;
                }
                run += ci;
            }
        }
    }
    __syncthreads();
}

// Block-wide radix select of local rank r among cand[0..c) (keys share all bits above their low `free_bits`),
// 8 bits at a time.  Returns the key, the number of candidates strictly below it and its multiplicity.
__device__ void block_radix_select(const unsigned long long* cand, unsigned long long c, unsigned long long r, int free_bits,
                                   unsigned int* sh_hist, unsigned long long* sh_io, unsigned long long& key_out,
                                   unsigned long long& below_out, unsigned long long& dup_out) {
    int remaining = free_bits;
    unsigned long long known = 0, mask = 0, below = 0, cnt = c;
    while (remaining > 0) {
        const int d = remaining >= 8 ? 8 : remaining;
        const int shift = remaining - d;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) sh_hist[i] = 0u;
        __syncthreads();
        for (unsigned long long i = threadIdx.x; i < c; i += blockDim.x) {
            const unsigned long long key = cand[i];
            if ((key & mask) == known) atomicAdd(&sh_hist[(unsigned)(key >> shift) & ((1u << d) - 1u)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp scan of the (up to) 256 bins, 8 per lane
            const int l = threadIdx.x;
            unsigned int cs[8];
            unsigned int mine = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { cs[i] = (l * 8 + i) < (1 << d) ? sh_hist[l * 8 + i] : 0u; mine += cs[i]; }
            unsigned int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(kFull, incl, o);
                if (l >= o) incl += v;
            }
            unsigned long long run = below + (unsigned long long)(incl - mine);
            // the bin holding rank r lies in exactly one lane's range (r < below + c always holds)
            if (r >= run && r < run + mine) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (r >= run && r < run + cs[i]) { sh_io[0] = run; sh_io[1] = (unsigned long long)(l * 8 + i); sh_io[2] = cs[i]; }
                    run += cs[i];
                }
            }
        }
        __syncthreads();
        below = sh_io[0];
        known |= sh_io[1] << shift;
        mask |= ((unsigned long long)((1u << d) - 1u)) << shift;
        cnt = sh_io[2];
        remaining = shift;
        __syncthreads();
    }
    key_out = (cand[0] & ~mask) | known;   // the bits above the free ones are common to all candidates
    below_out = below;
    dup_out = cnt;
}

struct FinishArgs {
    WalkState* st;
    unsigned long long* hist;             // [levels][4096], cleared here for the next step
    const unsigned long long* xch;        // sharded: all-reduced extras (kX*), else null -> read WalkState
    unsigned long long* blocks;           // candidate blocks: n_blocks of block_stride uint64 each
    int32_t n_blocks;
    long long block_stride;
    unsigned long long* own_block;        // this rank's block, reset for the next step
    unsigned long long* merged;           // scratch to concatenate blocks when n_blocks > 1 or for the slow path
    unsigned long long cand_cap;          // keys per block
    crwr_step_record* trace;
    int32_t do_select;                    // 0 on step 0 (utility.py:286)
    int32_t next_win;                     // 1: derive a select window for the next step from this step's cut (tail.cuh)
    int32_t dense_launched;               // 0: the step launched no large-row kernel, deferred rows are an error (host retries)
    float win_scale, win_min, win_max, win_rel_max;   // window half-width = clamp(win_scale x |relative move of the cut|, win_min, win_max)
    long long win_narrow_cands;           // sharded walk: candidates the ranks' push regions take together (with a margin)
};

// Shared-memory scratch the finish step needs (carved from the caller's buffer, 16-byte aligned).
constexpr size_t kFinishScratchBytes = sizeof(unsigned long long) * (kCandSmem + 4 + 32 + 4) + sizeof(unsigned int) * 256 +
                                       sizeof(int) * 4 + ((sizeof(WalkState) + 15) / 16) * 16 + 64;

// free_bits: low key bits in which the candidates may differ (-1: everything below the 24-bit prefix of the gather pass)
template <typename T>
__device__ __forceinline__ void finish_step_body(const FinishArgs& a, unsigned char* scratch, int free_bits = -1,
                                                 bool preloaded = false, bool hist_clean = false) {
    typedef KeyOf<T> KO;
    typedef typename KO::type K;
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(scratch);
    unsigned long long* sh_io = sk + kCandSmem;
    unsigned long long* wbest = sh_io + 4;
    unsigned long long* s64 = wbest + 32;                 // s_xk, s_xk1, s_total, s_succ
    unsigned int* sh_hist = reinterpret_cast<unsigned int*>(s64 + 4);
    int* s32 = reinterpret_cast<int*>(sh_hist + 256);     // s_have1, s_ovf
    WalkState* s_state_p = reinterpret_cast<WalkState*>(s32 + 4);
#define s_xk s64[0]
#define s_xk1 s64[1]
#define s_total s64[2]
#define s_succ s64[3]
#define s_have1 s32[0]
#define s_ovf s32[1]
    // the walk state is worked on in shared memory (one coalesced read, one coalesced write) instead of
    // through dozens of dependent L2 round trips by thread 0; no other CTA touches it at this point
    WalkState& s_state = *s_state_p;
    WalkState* const gst = a.st;
    const int t = threadIdx.x;
    if (!preloaded) {
        const unsigned* src = reinterpret_cast<const unsigned*>(gst);
        unsigned* dst = reinterpret_cast<unsigned*>(&s_state);
        for (int i = t; i < (int)(sizeof(WalkState) / 4); i += blockDim.x) dst[i] = __ldcg(src + i);
    }
    WalkState* st = &s_state;
    if (t == 0) { s_have1 = 0; s_xk = 0; s_xk1 = 0; s_ovf = 0; }
    __syncthreads();

    if (a.do_select) {
        // ---- concatenate the candidate blocks (one per rank)
        if (t == 0) {
            unsigned long long tot = 0, succ = ~0ull;
            int ovf = 0;
            for (int b = 0; b < a.n_blocks; ++b) {
                const unsigned long long* blk = a.blocks + (long long)b * a.block_stride;
                if (blk[0] > a.cand_cap) ovf = 1;
                tot += blk[0];
                succ = blk[1] < succ ? blk[1] : succ;
            }
            s_total = tot; s_succ = succ; s_ovf = ovf;
        }
        __syncthreads();
        const unsigned long long c = s_total;
        const unsigned long long r = st->rank_k - st->below;
        const unsigned long long* cand = a.blocks + kCandHeader;
        if (!s_ovf && a.n_blocks > 1) {
            unsigned long long off = 0;
            for (int b = 0; b < a.n_blocks; ++b) {
                const unsigned long long* blk = a.blocks + (long long)b * a.block_stride;
                const unsigned long long cb = blk[0];
                for (unsigned long long i = t; i < cb; i += blockDim.x) a.merged[off + i] = blk[kCandHeader + i];
                off += cb;
            }
            __syncthreads();
            cand = a.merged;
        }
        // a candidate set that fits shared memory is brought there once (the radix passes would read it three times)
        if (!s_ovf && c <= (unsigned long long)kCandSmem) {
            for (unsigned int i = t; i < c; i += blockDim.x) sk[i] = cand[i];
            __syncthreads();
            if (free_bits >= 0 && c > 512ull) cand = sk;
        }
        if (s_ovf) {
            // handled below (sticky error)
        } else if (c <= (unsigned long long)(free_bits >= 0 ? 512 : kCandSmem)) {
            for (unsigned int i = t; i < c; i += blockDim.x) {
                const unsigned long long me = sk[i];
                unsigned long long rank = 0;
                for (unsigned int j = 0; j < c; ++j) {
                    const unsigned long long o = sk[j];
                    rank += (o < me) || (o == me && j < i);
                }
                if (rank == r) s_xk = me;
                if (rank == r + 1) { s_xk1 = me; s_have1 = 1; }
            }
            __syncthreads();
        } else {
            unsigned long long key, below, dup;
            block_radix_select(cand, c, r, free_bits >= 0 ? free_bits : KO::bits - 2 * kSelectBits, sh_hist, sh_io, key, below, dup);
            if (r + 1 < below + dup) {
                if (t == 0) { s_xk = key; s_xk1 = key; s_have1 = 1; }
            } else {
                unsigned long long best = ~0ull;
                for (unsigned long long i = t; i < c; i += blockDim.x) {
                    const unsigned long long o = cand[i];
                    if (o > key && o < best) best = o;
                }
                best = warp_min(best);
                if ((t & 31) == 0) wbest[t >> 5] = best;
                __syncthreads();
                if (t == 0) {
                    unsigned long long b = ~0ull;
                    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b = wbest[w] < b ? wbest[w] : b;
                    s_xk = key;
                    if (below + dup < c) { s_xk1 = b; s_have1 = 1; }
                }
            }
            __syncthreads();
        }
    }

    if (t == 0) {
        double cut = __longlong_as_double(0x7ff8000000000000ll);   // NaN: no cut computed
        int cut_applied = 0;
        if (s_ovf) st->select_overflow = 1;
        if (a.do_select && !s_ovf) {
            const T xk = KO::dec((K)s_xk);
            T xk1 = xk;
            if (!st->clamp_last) xk1 = s_have1 ? KO::dec((K)s_xk1) : KO::dec((K)s_succ);
            // numpy _lerp
            const T g = (T)st->gamma;
            const T diff = sub_rn(xk1, xk);
            T out = add_rn(xk, mul_rn(diff, g));
            if (g >= (T)0.5) out = sub_rn(xk1, mul_rn(diff, sub_rn((T)1, g)));
            cut = (double)out;
            if (out > (T)0) cut_applied = 1;          // utility.py:288
        }
        unsigned long long l0, l1, l2, nnz_in, nnz_new, cap_ovf, sel_ovf, fb;
        if (a.xch) {
            l0 = a.xch[kXSq0]; l1 = a.xch[kXSq1]; l2 = a.xch[kXSq2];
            nnz_in = a.xch[kXNnzIn]; nnz_new = a.xch[kXNnzNew];
            cap_ovf = a.xch[kXCapOvf]; sel_ovf = a.xch[kXSelOvf]; fb = a.xch[kXFallback];
        } else {
            l0 = st->sq_limb[0]; l1 = st->sq_limb[1]; l2 = st->sq_limb[2];
            nnz_in = st->nnz_in; nnz_new = st->nnz_exact;
            cap_ovf = (unsigned long long)st->capacity_overflow; sel_ovf = 0; fb = st->fb_count;
        }
        if (sel_ovf) st->select_overflow = 1;
        if (cap_ovf) st->capacity_overflow = 1;
        const double sumsq = ((double)l0 * 0x1p-64 + (double)l1 * 0x1p-32) + (double)l2;
        const double delta = sqrt(sumsq);
        const int step = st->step;
        int done = 0, reason = CRWR_STOP_RUNNING;
        if (delta < st->tol) {                        // utility.py:293
            done = 1; reason = CRWR_STOP_TOLERANCE;
        } else {
            if (fabs(delta - st->delta_prev) < 0.001) st->plateau += 1;       // utility.py:298
            st->delta_prev = delta;                                           // utility.py:301
            if (st->plateau == 5) { done = 1; reason = CRWR_STOP_PLATEAU; }   // utility.py:302
        }
        if (step < kMaxTrace) {
            crwr_step_record rec;
            rec.nnz_in = (long long)nnz_in;
            rec.nnz_new = (long long)nnz_new;
            rec.cut = cut;
            rec.delta = delta;
            rec.cut_applied = cut_applied;
            rec.plateau_hits = st->plateau;
            rec.fallback_rows = (int)fb;
            rec.max_row_len = st->max_row_len;
            rec.max_kept = st->max_kept;
            rec.fast_rows = (int)st->fast_rows;
            a.trace[step] = rec;
        }
        // deferred rows with no large-row kernel behind the propagation: the step is incomplete, the host retries
        if (fb > 0 && !a.dense_launched) st->fb_unhandled = 1;
        // select window of the next step: the cut moves less and less, so it will lie within a few times its last move
        {
            int on = 0, narrow = 0;
            const bool moved = cut_applied && st->cut_on && st->cut > 0.0 && cut > 0.0;
            const double rel = moved ? fabs(cut - st->cut) / cut : 0.0;
            if (a.next_win && moved) {
                const double c1 = cut;
                // a settling cut does not move monotonically less: a step of 0.06 % can be followed by one of 0.4 %.  Half
                // the move before the last is the floor (a miss costs a one-CTA select over the whole value array)
                const double rel_w = rel > 0.5 * st->win_prev_rel ? rel : 0.5 * st->win_prev_rel;
                double dl = (double)a.win_scale * rel_w;
                dl = dl < (double)a.win_min ? (double)a.win_min : (dl > (double)a.win_max ? (double)a.win_max : dl);
                const T lo = (T)(c1 * (1.0 - dl)), hi = (T)(c1 * (1.0 + dl));
                if (rel <= (double)a.win_rel_max && lo > (T)0 && hi > lo) {
                    on = 1;
                    // sharded walk: about 0.4 x dl of the values fall into the window (measured on the synthetic workloads,
                    // whose value density at the cut is ~0.2 per unit of log value); they must fit the ranks' push regions
                    narrow = 0.4 * dl * (double)nnz_new <= (double)a.win_narrow_cands ? 1 : 0;
                    st->win_lo = (unsigned long long)KO::enc(lo);
                    st->win_hi = (unsigned long long)KO::enc(hi);
                }
            }
            st->win_on = on;
            st->win_narrow = narrow;
            st->win_prev_rel = rel;
        }
        // rows written in slot-claim order (row-group kernel) are only valid operands behind a cut
        if (st->unordered_rows && a.do_select && !cut_applied && st->n_total > 0 && !st->capacity_overflow && !st->select_overflow)
            st->order_hazard = 1;
        st->unordered_rows = 0;
        st->last_delta = delta;
        st->last_nnz_new = nnz_new;
        st->cut_on = cut_applied;
        if (cut_applied) st->cut = cut;
        st->parity ^= 1;
        st->step = step + 1;
        if (!done && step + 1 >= st->max_steps) { done = 1; reason = CRWR_STOP_MAX_STEPS; }
        if (st->capacity_overflow || st->select_overflow || st->peer_timeout || st->order_hazard || st->fb_unhandled) done = 1;
        st->done = done;
        st->stop_reason = reason;
        st->last_fallback_rows = (int)fb;
        st->last_max_kept = st->max_kept;
        st->last_max_row_len = st->max_row_len;
        st->max_kept = 0;
        if (!done) {
            // counters of the next step (when stopping, cursor/max_row_len keep describing the result)
            st->cursor = 0;
            st->nnz_exact = 0;
            st->max_row_len = 0;
        }
        st->nnz_in = 0;
        st->row_counter = 0;
        st->fb_counter = 0;
        st->fb_count = 0;
        st->below = 0;
        st->prefix = 0;
        st->sq_limb[0] = 0; st->sq_limb[1] = 0; st->sq_limb[2] = 0;
        st->tick_dense = 0; st->tick_hist1 = 0; st->tick_gather = 0; st->tick_prop = 0;
        st->scan_done = 0;
        st->hist1_ready = 0;
        st->slow_count = 0; st->slow2_count = 0; st->fast_rows = 0; st->built_rows = 0; st->win_below = 0;
        if (cut_applied) {      // the next cut most likely falls into the same level-0 bin
            st->pred_prefix0 = (unsigned long long)(KO::enc((T)cut) >> (KO::bits - kSelectBits));
            st->pred_valid = 1;
        } else {
            st->pred_valid = 0;
        }
        a.own_block[0] = 0;
        a.own_block[1] = ~0ull;
    }
    __syncthreads();   // thread 0 has consumed the exchanged counters and updated the state
    {
        const unsigned* src = reinterpret_cast<const unsigned*>(&s_state);
        unsigned* dst = reinterpret_cast<unsigned*>(gst);
        for (int i = t; i < (int)(sizeof(WalkState) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (!hist_clean) for (int i = t; i < kSelectLevels * kSelectBins + kXExtras; i += blockDim.x) a.hist[i] = 0ull;
}

// Where finish_step_body keeps the walk state inside its scratch (for a caller that loads it beforehand).
__device__ __forceinline__ WalkState* finish_state_ptr(unsigned char* scratch) {
    unsigned long long* sk = reinterpret_cast<unsigned long long*>(scratch);
    unsigned int* sh_hist = reinterpret_cast<unsigned int*>(sk + kCandSmem + 4 + 32 + 4);
    return reinterpret_cast<WalkState*>(reinterpret_cast<int*>(sh_hist + 256) + 4);
}

#undef s_xk
#undef s_xk1
#undef s_total
#undef s_succ
#undef s_have1
#undef s_ovf

// One histogram pass over the value array by the whole grid: level 0 counts the top 12 key bits of every
// value, level 1 the next 12 bits of the values inside level-0 bin `prefix`.  `sh` is a 4096-bin
// shared-memory histogram private to the CTA; it is flushed with one atomic per non-empty bin.
// The value array of a step is two ranges: [0, n) the per-step region, [fixed_base, fixed_base + n_fixed) the fixed slots
// of rows with a cached structure (structure.cuh); entry i >= n of the concatenation lives at fixed_base + (i - n).
struct ValRanges { long long n, fixed_base, n_fixed; };
__device__ __forceinline__ ValRanges val_ranges(const WalkState* st, long long fixed_base) {
    ValRanges r;
    r.n = (long long)st->cursor;
    r.fixed_base = fixed_base;
    r.n_fixed = (long long)st->slot_cursor;
    if (r.n_fixed > fixed_base - 16) r.n_fixed = fixed_base - 16;   // the fixed region is at least as long as that (api.cu, make_layout)
    return r;
}
template <typename T>
__device__ __forceinline__ T range_load(const T* __restrict__ vals, const ValRanges& r, long long i) {
    if (i < r.n) return vals[i];
    if (i < r.n + r.n_fixed) return vals[r.fixed_base + (i - r.n)];
    return (T)0;
}

template <typename T>
__device__ __forceinline__ void hist_pass(const T* __restrict__ vals, const ValRanges rg, typename KeyOf<T>::type prefix, int level,
                                          unsigned int* sh, unsigned long long* hist) {
    const long long n = rg.n + rg.n_fixed;
    typedef KeyOf<T> KO;
    typedef typename KO::type K;
    for (int i = threadIdx.x; i < kSelectBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        T vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) vv[u] = range_load(vals, rg, i0 + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T v = vv[u];
            if (v == (T)0) continue;                  // zero padding of a row reservation (or past the end)
            const K key = KO::enc(v);
            if (level == 0) {
                atomicAdd(&sh[(unsigned)(key >> (KO::bits - kSelectBits))], 1u);
            } else {
                if ((key >> (KO::bits - kSelectBits)) == prefix)
                    atomicAdd(&sh[(unsigned)(key >> (KO::bits - 2 * kSelectBits)) & (kSelectBins - 1)], 1u);
            }
        }
    }
    __syncthreads();
    unsigned long long* out = hist + (size_t)level * kSelectBins;
    for (int i = threadIdx.x; i < kSelectBins; i += blockDim.x) {
        unsigned int c = sh[i];
        if (c) atomicAdd(&out[i], (unsigned long long)c);
    }
}

template <typename T>
__global__ void __launch_bounds__(512) select_hist_kernel(const T* vals, long long fixed_base, WalkState* st,
                                                          unsigned long long* __restrict__ hist, int level, int fused_tail) {
    __shared__ unsigned int sh[kSelectBins];
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    if (level == 1 && st->hist1_ready) return;      // the propagate epilogue already filled this level and it was scanned
    hist_pass<T>(vals, val_ranges(st, fixed_base), (typename KeyOf<T>::type)st->prefix, level, sh, hist);
    if (fused_tail && last_block(&st->tick_hist1)) scan_level<T, 512>(st, hist, level);
}

template <typename T>
__global__ void __launch_bounds__(1024) select_scan_kernel(WalkState* st, const unsigned long long* __restrict__ hist,
                                                           int level) {
    if (st->done) return;
    scan_level<T, 1024>(st, hist, level);
}

// Candidate block layout (uint64): [0] count, [1] smallest key above the candidate bin, [2..] keys.
// One pass over the value array by the whole grid.
template <typename T>
__device__ __forceinline__ void gather_pass(const T* __restrict__ vals, const ValRanges rg, typename KeyOf<T>::type prefix,
                                            unsigned long long* block, unsigned long long cand_cap) {
    typedef KeyOf<T> KO;
    typedef typename KO::type K;
    const long long n = rg.n + rg.n_fixed;
    const long long stride = (long long)gridDim.x * blockDim.x;
    K best = ~(K)0;
    bool have = false;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        T vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) vv[u] = range_load(vals, rg, i0 + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T v = vv[u];
            if (v == (T)0) continue;                  // zero padding of a row reservation (or past the end)
            const K key = KO::enc(v);
            const K p = key >> (KO::bits - 2 * kSelectBits);
            if (p == prefix) {
                unsigned long long idx = atomicAdd(&block[0], 1ull);
                if (idx < cand_cap) block[kCandHeader + idx] = (unsigned long long)key;
            } else if (p > prefix) {
                if (!have || key < best) best = key;
                have = true;
            }
        }
    }
    unsigned long long b64 = have ? (unsigned long long)best : ~0ull;
    b64 = warp_min(b64);
    if ((threadIdx.x & 31) == 0 && b64 != ~0ull) atomicMin(&block[1], b64);
}

// fused (unsharded walks): the last CTA to finish runs the step's finish logic.
template <typename T>
__global__ void __launch_bounds__(512) select_gather_kernel(const T* vals, long long fixed_base, WalkState* st,
                                                            unsigned long long* __restrict__ block, unsigned long long cand_cap,
                                                            int fused, FinishArgs fin) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
    pdl_wait();
    pdl_launch_dependents();
    if (st->done) return;
    gather_pass<T>(vals, val_ranges(st, fixed_base), (typename KeyOf<T>::type)st->prefix, block, cand_cap);
    if (fused && last_block(&st->tick_gather)) finish_step_body<T>(fin, scratch);
}

template <typename T>
__global__ void __launch_bounds__(1024) finish_step_kernel(FinishArgs a) {
    __shared__ __align__(16) unsigned char scratch[kFinishScratchBytes];
    pdl_wait();
    pdl_launch_dependents();
    if (a.st->done) return;
    finish_step_body<T>(a, scratch);
}

// Sharded walk: publish this rank's step counters behind the level-0 histogram so that ONE
// all-reduce (uint64 SUM) carries histogram, counts and the fixed-point delta^2.
__global__ void pack_exchange_kernel(WalkState* st, unsigned long long* extras) {
    if (st->done) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        extras[kXSq0] = st->sq_limb[0];
        extras[kXSq1] = st->sq_limb[1];
        extras[kXSq2] = st->sq_limb[2];
        extras[kXNnzIn] = st->nnz_in;
        extras[kXNnzNew] = st->nnz_exact;
        extras[kXCapOvf] = (unsigned long long)st->capacity_overflow;
        extras[kXSelOvf] = (unsigned long long)st->select_overflow;
        extras[kXFallback] = st->fb_count;
    }
}

}  // namespace crwr
