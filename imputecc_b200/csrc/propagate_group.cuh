// K2, row-group form: the same walk-step propagation as propagate.cuh (utility.py:282-283 fused with the
// lazy cut of utility.py:290 and the level-0/1 histograms of utility.py:287), re-cut for short rows and
// for steps whose output is consumed AFTER a percentile cut (every step but step 0).
//
// Work decomposition: G lanes (4, 8, 16 or 32) own one restart row, so a warp carries 32/G rows at once
// and every warp instruction of the per-row bookkeeping (row read, cut, sort, epilogue, write) is
// shared by 32/G rows.  The G lanes consume the row's products ONE P ROW PER ROUND: the entries of one
// row of P have distinct columns, so the lanes of a round never meet in an accumulator slot, and the
// rounds follow each other in the order of the kept operand entries - every output is therefore summed
// in exactly scipy's order (k ascending after a cut; separate multiply and add) without any duplicate
// detection between lanes.  P is read straight from global memory (it stays L2 resident; the next
// round's entries are requested before the current round is accumulated).
//
// Accumulators: a per-row hash table in shared memory made of two-slot buckets.  A round probes exactly
// one bucket (one 64-bit load of both keys, no probe loop, no divergence): hit -> add; free slot ->
// claim; bucket full of other columns -> the product is appended, in order, to a small overflow list
// that one lane folds in by linear probing after the last round (a column either lives in its bucket
// or has ALL its products in the list, so the order of its additions is kept either way).
//
// Storage order: scipy stores a product row in first-touch order, which matters only to a consumer that
// reads the row WITHOUT a cut (the cut sorts the operand by column).  Every stored value is positive, so
// the percentile of steps >= 1 is positive and the cut is always applied: only step 0's output is ever
// read in storage order, and step 0 runs on the warp-per-row kernel of propagate.cuh, which keeps it.
// Here rows are written in slot-claim order; the finish step raises `order_hazard` (a hard error) should
// a step >= 1 ever end without a cut although values were stored.
//
// Rows whose kept list, output or overflow list does not fit are deferred to the large-row kernel.
#pragma once
#include "structure.cuh"

namespace crwr {

struct GroupShape {
    int H;              // accumulator slots per row (even: two-slot buckets; >= kmax)
    int limit;          // max distinct output columns held (< H)
    int ov_cap;         // overflow list entries per row
    int kmax;           // max kept operand entries
    int G;              // lanes per row
    int warps;          // warps per CTA
    int rank_sort_max;  // kept lists up to this length are sorted by rank counting, longer ones by a bitonic network
    size_t per_row;     // shared memory per row
};

template <typename T>
__host__ __device__ inline size_t group_smem_per_row(int H, int limit, int kmax, int ov_cap) {
    return align16((size_t)H * sizeof(T)) + align16((size_t)kmax * sizeof(T)) + align16((size_t)ov_cap * sizeof(T)) +
           align16((size_t)H * 4) + 3 * align16((size_t)kmax * 4) + align16((size_t)ov_cap * 4) + align16((size_t)limit * 2);
}

__device__ __forceinline__ int hash_bucket2(int32_t j, int n_buckets) {
    return (int)(__umulhi((uint32_t)j * 0x9E3779B1u, (uint32_t)n_buckets) << 1);     // first slot of the bucket
}

// Ascending in-place sort of (key[i], val[i]), i < n, by the G lanes of a group; the other groups of the
// warp run the same stages on their own arrays (nmax = longest list in the warp bounds the stage loops;
// the extra stages of a shorter list compare sorted data and change nothing).  Keys are unique.
template <typename T, int G>
__device__ __forceinline__ void group_sort_pairs(int32_t* key, T* val, int n, int nmax, int sub) {
    for (int k = 2; (k >> 1) < nmax; k <<= 1) {
        for (int i = sub; i < n; i += G) {
            const int l = i ^ (k - 1);
            if (l > i && l < n) {
                const int32_t x = key[i], y = key[l];
                if (y < x) { key[i] = y; key[l] = x; const T t = val[i]; val[i] = val[l]; val[l] = t; }
            }
        }
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = sub; i < n; i += G) {
                const int l = i ^ j;
                if (l > i && l < n) {
                    const int32_t x = key[i], y = key[l];
                    if (y < x) { key[i] = y; key[l] = x; const T t = val[i]; val[i] = val[l]; val[l] = t; }
                }
            }
            __syncwarp();
        }
    }
}

// Longest rows first: a one-off counting sort of the rows by the length of their current iterate row, run once the
// walk has settled (the lengths barely move afterwards).  The row counter then hands out the heavy rows first and the
// lightest last, so the kernel does not end on a long row that started late, and the rows sharing a warp are of
// similar weight.  Order inside a length bucket is arbitrary: the order of processing never affects a value.
__global__ void __launch_bounds__(1024) row_order_kernel(const int32_t* __restrict__ row_len, int n_rows, int32_t* __restrict__ order) {
    constexpr int NB = 2048;
    __shared__ int cnt[NB];
    __shared__ int warp_tot[32];
    const int t = threadIdx.x;
    for (int i = t; i < NB; i += 1024) cnt[i] = 0;
    __syncthreads();
    // bucket 0 = longest
    for (int r = t; r < n_rows; r += 1024) {
        int len = row_len[r];
        len = len < 0 ? 0 : (len > NB - 1 ? NB - 1 : len);
        atomicAdd(&cnt[NB - 1 - len], 1);
    }
    __syncthreads();
    // exclusive scan of the 2048 counts: 2 per thread
    const int a0 = cnt[2 * t], a1 = cnt[2 * t + 1];
    int incl = a0 + a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        const int w = warp_tot[t];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, wi, o);
            if (t >= o) wi += v;
        }
        warp_tot[t] = wi - w;
    }
    __syncthreads();
    const int excl = warp_tot[t >> 5] + incl - (a0 + a1);
    cnt[2 * t] = excl;
    cnt[2 * t + 1] = excl + a0;
    __syncthreads();
    for (int r = t; r < n_rows; r += 1024) {
        int len = row_len[r];
        len = len < 0 ? 0 : (len > NB - 1 ? NB - 1 : len);
        order[atomicAdd(&cnt[NB - 1 - len], 1)] = r;
    }
}

template <typename T, int G>
__global__ void __launch_bounds__(256) propagate_group_kernel(PropArgs<T> a, GroupShape sh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RPW = 32 / G;                      // rows per warp
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int sub = lane & (G - 1);
    const int grp = lane / G;
    const unsigned gmask = G == 32 ? kFull : (((1u << G) - 1u) << (grp * G));
    const unsigned glt = gmask & lanemask_lt();      // lanes of my group below me
    const int H = sh.H, KMAX = sh.kmax, LIMIT = sh.limit, OV = sh.ov_cap;
    const int NB = H >> 1;

    unsigned char* p = smem_raw + (size_t)(warp * RPW + grp) * sh.per_row;
    T* tval = reinterpret_cast<T*>(p);             p += align16((size_t)H * sizeof(T));
    T* kval = reinterpret_cast<T*>(p);             p += align16((size_t)KMAX * sizeof(T));
    T* oval = reinterpret_cast<T*>(p);             p += align16((size_t)OV * sizeof(T));
    int32_t* tkey = reinterpret_cast<int32_t*>(p); p += align16((size_t)H * 4);
    int32_t* kcol = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* kps = reinterpret_cast<int32_t*>(p);  p += align16((size_t)KMAX * 4);
    int32_t* klen = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* ocol = reinterpret_cast<int32_t*>(p); p += align16((size_t)OV * 4);
    uint16_t* ord = reinterpret_cast<uint16_t*>(p);

    // slot of column j, or -1: its bucket first, then (bucket full of other columns) the linear probe
    // sequence the overflow list was folded in with
    auto lookup = [&](int32_t j) -> int {
        int s = hash_bucket2(j, NB);
        const int2 kk = *reinterpret_cast<const int2*>(tkey + s);
        if (kk.x == j) return s;
        if (kk.y == j) return s + 1;
        if (kk.x == -1 || kk.y == -1) return -1;
        s += 2;
        if (s == H) s = 0;
        for (;;) {
            const int32_t k = tkey[s];
            if (k == j) return s;
            if (k == -1) return -1;
            if (++s == H) s = 0;
        }
    };

    unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw + (size_t)sh.warps * RPW * sh.per_row);
    __shared__ unsigned int s_wn;
    if (threadIdx.x == 0) s_wn = 0u;
    const WinStage stage = {reinterpret_cast<unsigned long long*>(win), &s_wn};     // window mode reuses the histogram window
    WinAcc wacc = {0ull, ~0ull};
    for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) win[i] = 0u;
    for (int i = sub; i < H; i += G) tkey[i] = -1;
    pdl_wait();                                      // the tables above are set up while the previous kernel drains
    if (a.st->done) return;
    __syncthreads();

    if (blockIdx.x == 0 && threadIdx.x == 0) a.st->unordered_rows = 1;
    const bool cut_on = __ldcg(&a.st->cut_on) != 0;
    const T cut = (T)__ldcg(&a.st->cut);
    const bool win_mode = a.win_mode && __ldcg(&a.st->win_on) != 0;
    const unsigned long long wlo = __ldcg(&a.st->win_lo), whi = __ldcg(&a.st->win_hi);
    const int n_rows = a.row_list_count ? (int)__ldcg(a.row_list_count) : a.n_rows;
    const bool pred_valid = a.prefill_l1 && __ldcg(&a.st->pred_valid) != 0;
    const typename KeyOf<T>::type pred0 = (typename KeyOf<T>::type)__ldcg(&a.st->pred_prefix0);
    const int32_t* __restrict__ pcols = a.p_cols;
    const T* __restrict__ pvals = a.p_vals;

    // per-lane totals: only the first lane of a group records its row; reduced over the warp at the end
    unsigned long long t_sq0 = 0, t_sq1 = 0, t_sq2 = 0, t_nnz = 0, t_kept = 0;
    int t_max_kept = 0, t_max_len = 0;

    // Rows are handed out by an atomic counter.  (Measured and rejected: a static first pass - 9 % slower on long
    // rows, where neighbouring rows of one genome end up in one CTA - and asking for the next row at the start
    // of a pass - 7-10 % slower.)
    // a slow list shorter than the grid: the surplus warps leave without touching the row counter (thousands of
    // same-address atomics would otherwise cost more than the few listed rows)
    if (a.row_list_count && (int)((blockIdx.x * (blockDim.x >> 5) + warp) * RPW) >= n_rows) goto rows_done;
    for (;;) {
        int r0 = 0;
        if (lane == 0) r0 = (int)atomicAdd(&a.st->row_counter, (unsigned)RPW);
        r0 = __shfl_sync(kFull, r0, 0);
        if (r0 >= n_rows) break;
        const bool has_row = r0 + grp < n_rows;
        const int row = (has_row && a.row_order) ? __ldg(a.row_order + r0 + grp) : r0 + grp;

        // ---- read the row, dropping entries at or below the previous cut (strict >, utility.py:290).
        // With a cut the kept entries must be consumed in column order (`Q > cut` sorts the operand,
        // scipy _compressed.py:_scalar_binopt): they are parked in kps / tval and sorted into kcol / kval.
        long long ptr = 0;
        int len = 0;
        if (has_row) { ptr = a.in.row_ptr[row]; len = a.in.row_len[row]; }
        const int32_t* rc = a.in.cols + ptr;
        const T* rv = a.in.vals + ptr;
        const int maxlen = __reduce_max_sync(kFull, len);
        int32_t* ucol = cut_on ? kps : kcol;
        T* uval = cut_on ? tval : kval;
        int nk = 0;                   // kept entries
        for (int b = 0; b < maxlen; b += 4 * G) {
            int32_t c[4];
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * G + sub;
                c[u] = 0; v[u] = (T)0;
                if (e < len) { c[u] = __ldg(rc + e); v[u] = __ldg(rv + e); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * G + sub;
                // a stored value is never zero: zeros are fixed-slot entries that count as absent (structure.cuh)
                const bool keep = (e < len) && (cut_on ? v[u] > cut : v[u] != (T)0);
                const unsigned m = __ballot_sync(kFull, keep) & gmask;
                const int pos = nk + __popc(m & glt);
                if (keep && pos < KMAX) { CRWR_CHECK(a.st, c[u] >= 0, 4); ucol[pos] = c[u]; uval[pos] = v[u]; }
                nk += __popc(m);
            }
        }
        __syncwarp();
        t_max_kept = max(t_max_kept, nk);
        bool defer = nk > KMAX;
        int nk_live = (has_row && !defer) ? nk : 0;
        const int maxnk = __reduce_max_sync(kFull, nk_live);

        if (cut_on) {
            if (maxnk <= sh.rank_sort_max) {
                for (int b = 0; b < maxnk; b += G) {
                    const int e = b + sub;
                    if (e < nk_live) {
                        const int32_t c = kps[e];
                        const T v = tval[e];
                        int rank = 0;
#pragma unroll 4
                        for (int i = 0; i < nk_live; ++i) rank += (kps[i] < c) ? 1 : 0;
                        kcol[rank] = c;
                        kval[rank] = v;
                    }
                }
            } else {
                group_sort_pairs<T, G>(kps, tval, nk_live, maxnk, sub);
                for (int e = sub; e < nk_live; e += G) { kcol[e] = kps[e]; kval[e] = tval[e]; }
            }
            __syncwarp();
        }
        // extent of the P row of every kept entry
        for (int b = 0; b < maxnk; b += G) {
            const int idx = b + sub;
            if (idx < nk_live) {
                const int k = kcol[idx];
                CRWR_CHECK(a.st, k >= 0, 5);
                const int ps = __ldg(a.p_indptr + k);
                kps[idx] = ps;
                klen[idx] = __ldg(a.p_indptr + k + 1) - ps;
            }
        }
        __syncwarp();

        // ---- the row's products, one P row (G entries of it) per round; the entries of round r+1 are
        // requested from global memory before round r is accumulated
        int count = 0, nov = 0;
        int c_idx = 0, c_e = 0, c_pe = 0;
        T c_q = (T)0;
        if (nk_live > 0) { c_e = kps[0]; c_pe = c_e + klen[0]; c_q = kval[0]; }
        int32_t j_n = -1;
        T w_n = (T)0, q_n = (T)0;
        bool act_n = false;
        auto fetch = [&]() {
            act_n = c_idx < nk_live;
            const int my = c_e + sub;
            j_n = -1; w_n = (T)0; q_n = c_q;
            if (act_n && my < c_pe) { j_n = __ldg(pcols + my); w_n = __ldg(pvals + my); }
            if (act_n) {
                c_e += G;
                if (c_e >= c_pe) {
                    ++c_idx;
                    if (c_idx < nk_live) { c_e = kps[c_idx]; c_pe = c_e + klen[c_idx]; c_q = kval[c_idx]; }
                }
            }
        };
        fetch();
        while (__any_sync(kFull, act_n)) {
            const int32_t j = j_n;
            const T prod = mul_rn(q_n, w_n);
            const bool over = act_n && (count + nov + G > LIMIT);    // the table may not take another round: defer the row
            if (over) { defer = true; nk_live = 0; }
            fetch();
            bool is_new = false, ovf = false;
            int slot = 0;
            if (j >= 0 && !over) {
                const int b2 = hash_bucket2(j, NB);
                const int2 kk = *reinterpret_cast<const int2*>(tkey + b2);
                if (kk.x == j) tval[b2] = add_rn(tval[b2], prod);
                else if (kk.y == j) tval[b2 + 1] = add_rn(tval[b2 + 1], prod);
                else if (kk.x == -1 && atomicCAS(&tkey[b2], -1, j) == -1) { slot = b2; is_new = true; }
                else if (kk.y == -1 && atomicCAS(&tkey[b2 + 1], -1, j) == -1) { slot = b2 + 1; is_new = true; }
                else ovf = true;       // bucket full of other columns (or both claims lost to lanes of this round)
                if (is_new) tval[slot] = prod;
            }
            const unsigned nm = __ballot_sync(kFull, is_new) & gmask;
            const unsigned om = __ballot_sync(kFull, ovf) & gmask;
            if (is_new) { CRWR_CHECK(a.st, count + __popc(nm & glt) < LIMIT, 2); ord[count + __popc(nm & glt)] = (uint16_t)slot; }
            count += __popc(nm);
            if (ovf) {
                const int pos = nov + __popc(om & glt);
                if (pos < OV) { ocol[pos] = j; oval[pos] = prod; }
            }
            nov += __popc(om);
            __syncwarp();
        }
        if (nov > OV) defer = true;
        // fold the overflow list in, in order, by linear probing from the home bucket (one lane per row)
        if (has_row && !defer && nov > 0 && sub == 0) {
            for (int t = 0; t < nov; ++t) {
                const int32_t j = ocol[t];
                const T prod = oval[t];
                int s_ = hash_bucket2(j, NB);
                for (;;) {
                    const int32_t kk = tkey[s_];
                    if (kk == j) { tval[s_] = add_rn(tval[s_], prod); break; }
                    if (kk == -1) { tkey[s_] = j; tval[s_] = prod; ord[count++] = (uint16_t)s_; break; }
                    if (++s_ == H) s_ = 0;
                }
            }
        }
        count = __shfl_sync(kFull, count, grp * G);
        __syncwarp();

        if (has_row && defer) {
            // undo and hand the row to the large-row kernel
            for (int i = sub; i < count; i += G) tkey[ord[i]] = -1;
            if (sub == 0) a.fb_rows[atomicAdd(&a.st->fb_count, 1u)] = row;
        }
        const bool live = has_row && !defer;
        const int cnt = live ? count : 0;
        nk_live = live ? nk : 0;

        // ---- epilogue: scale, restart add, norm, histogram, write
        const int32_t dcol = a.row_begin + row;
        int dslot = -1;
        if (live) dslot = lookup(dcol);
        const bool diag_new = live && dslot < 0;
        const int n_cand = cnt + (diag_new ? 1 : 0);
        // one reservation of the product buffer per warp: exact size, no padding between rows
        int offs = 0, total = 0;
#pragma unroll
        for (int g = 0; g < RPW; ++g) {
            const int v = __shfl_sync(kFull, n_cand, g * G);
            if (g < grp) offs += v;
            total += v;
        }
        long long base = 0;
        if (lane == 0 && total > 0) base = (long long)atomicAdd(&a.st->cursor, (unsigned long long)total);

        const int maxcnt = __reduce_max_sync(kFull, cnt);
        Sq128 pos = {0ull, 0ull}, neg = {0ull, 0ull};
        int nz = 0;
        for (int b = 0; b < maxcnt; b += G) {
            const int i = b + sub;
            bool one = false;
            T v = (T)0;
            if (i < cnt) {
                const int slot = ord[i];
                v = mul_rn(a.scale, tval[slot]);
                if (slot == dslot) v = add_rn(v, a.restart);
                tval[slot] = v;
                sq_add(pos, (double)v * (double)v);
                one = v != (T)0;
            }
            nz += __popc(__ballot_sync(kFull, one) & gmask);
            if (win_mode) win_classify<T>(one, v, wlo, whi, wacc, stage, a.cand, a.cand_cap);
            else if (a.hist) hist_values<T>(a.hist, win, one, v, pred_valid, pred0);
        }
        if (diag_new) {
            if (sub == 0) sq_add(pos, (double)a.restart * (double)a.restart);
            if (a.restart != (T)0) nz += 1;
        }
        if (win_mode) win_classify<T>(diag_new && sub == 0 && a.restart != (T)0, a.restart, wlo, whi, wacc, stage, a.cand, a.cand_cap);
        else if (a.hist && __any_sync(kFull, diag_new))
            hist_values<T>(a.hist, win, diag_new && sub == 0 && a.restart != (T)0, a.restart, pred_valid, pred0);
        __syncwarp();
        for (int b = 0; b < maxnk; b += G) {
            const int idx = b + sub;
            if (idx < nk_live) {
                const int32_t j = kcol[idx];
                const T oldv = kval[idx];
                const int slot = lookup(j);
                const bool found = slot >= 0;
                T newv = (T)0;
                if (found) newv = tval[slot];
                else if (j == dcol) newv = a.restart;
                const T d = sub_rn(oldv, newv);
                sq_add(pos, (double)d * (double)d);
                sq_add(neg, (double)newv * (double)newv);
            }
        }
        const Sq128 rowsq = sq_diff(sq_group_sum(pos, G), sq_group_sum(neg, G));
        if (live && sub == 0) {
            t_sq2 += rowsq.hi;
            t_sq1 += rowsq.lo >> 32;
            t_sq0 += rowsq.lo & 0xffffffffull;
            t_kept += (unsigned long long)nk;
            t_nnz += (unsigned long long)nz;
            t_max_len = max(t_max_len, nz);
        }

        base = __shfl_sync(kFull, base, 0);
        const bool fits = base + total <= a.out_cap;
        if (!fits && lane == 0) a.st->capacity_overflow = 1;
        base += offs;
        if (live && sub == 0) {
            a.out.row_ptr[row] = fits ? base : 0;
            a.out.row_len[row] = fits ? nz : 0;
        }
        if (fits) {
            const int maxcand = __reduce_max_sync(kFull, n_cand);
            const int shift = diag_new ? 1 : 0;
            int written = 0;
            for (int b = 0; b < maxcand; b += G) {
                const int i = b + sub;
                int32_t c = 0;
                T v = (T)0;
                if (i < n_cand) {
                    if (i < shift) { c = dcol; v = a.restart; }
                    else { const int slot = ord[i - shift]; c = tkey[slot]; v = tval[slot]; }
                }
                const bool nzv = (i < n_cand) && (v != (T)0);
                const unsigned m = __ballot_sync(kFull, nzv) & gmask;
                if (nzv) {
                    const long long o = base + written + __popc(m & glt);
                    CRWR_CHECK(a.st, o >= 0 && o < a.out_cap, 6);
                    a.out.cols[o] = c;
                    a.out.vals[o] = v;
                }
                written += __popc(m);
            }
            // zero values are never stored (csr_matmat / csr_plus_csr): the unused tail is zero-padded
            for (int i = written + sub; i < n_cand; i += G) { a.out.cols[base + i] = -1; a.out.vals[base + i] = (T)0; }
        }
        __syncwarp();
        for (int i = sub; i < cnt; i += G) tkey[ord[i]] = -1;
        __syncwarp();
    }
rows_done:
    pdl_launch_dependents();        // this warp is out of rows: the next kernel's CTAs may take the SM as it empties

    // ---- publish the warp's totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t_sq0 += __shfl_xor_sync(kFull, t_sq0, o);
        t_sq1 += __shfl_xor_sync(kFull, t_sq1, o);
        t_sq2 += __shfl_xor_sync(kFull, t_sq2, o);
        t_nnz += __shfl_xor_sync(kFull, t_nnz, o);
        t_kept += __shfl_xor_sync(kFull, t_kept, o);
    }
    t_max_kept = __reduce_max_sync(kFull, t_max_kept);
    t_max_len = __reduce_max_sync(kFull, t_max_len);
    // one set of global atomics per CTA: the warps meet in shared memory first
    __shared__ unsigned long long s_tot[5];
    __shared__ int s_max[2];
    if (threadIdx.x < 5) s_tot[threadIdx.x] = 0ull;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) {
        if (t_sq0) atomicAdd(&s_tot[0], t_sq0);
        if (t_sq1) atomicAdd(&s_tot[1], t_sq1);
        if (t_sq2) atomicAdd(&s_tot[2], t_sq2);
        if (t_nnz) atomicAdd(&s_tot[3], t_nnz);
        if (t_kept) atomicAdd(&s_tot[4], t_kept);
        if (t_max_kept) atomicMax(&s_max[0], t_max_kept);
        if (t_max_len) atomicMax(&s_max[1], t_max_len);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        WalkState* st = a.st;
        if (s_tot[0]) atomicAdd(&st->sq_limb[0], s_tot[0]);
        if (s_tot[1]) atomicAdd(&st->sq_limb[1], s_tot[1]);
        if (s_tot[2]) atomicAdd(&st->sq_limb[2], s_tot[2]);
        if (s_tot[3]) atomicAdd(&st->nnz_exact, s_tot[3]);
        if (s_tot[4]) atomicAdd(&st->nnz_in, s_tot[4]);
        if (s_max[0]) atomicMax(&st->max_kept, s_max[0]);
        if (s_max[1]) atomicMax(&st->max_row_len, s_max[1]);
    }
    if (win_mode) win_flush(wacc, stage, a.cand, a.cand_cap, a.st);
    else if (a.hist) hist_window_flush<T>(a.hist, win);
}

}  // namespace crwr
