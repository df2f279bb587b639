// K2, cached-structure form: the walk step's propagation (utility.py:282-283 fused with the lazy cut of
// utility.py:290) for rows whose SYMBOLIC work is already known.
//
// From the third walk step on the kept set of a restart row barely moves (the walk plateaus),
// yet the row-group kernel redoes the whole symbolic part of the row's SpGEMM every step: hashing, slot claims,
// first-touch lists.  Here the symbolic result is cached per row, once, for a SUPERSET K+ of the kept columns
// (the row's entries above alpha x cut, alpha < 1):
//     K+            the operand columns, ascending          O             the output columns they reach (+ restart column)
//     per output    its contributions (index into K+, P value) in ascending k = exactly scipy's summation order
// A later step may use the structure whenever its kept set is INSIDE K+: members of K+ at or below the cut take
// q = 0, their products are exact zeros and adding them changes no bit; outputs reached only through them come
// out 0 and are dropped like csr_matmat drops them.  The numeric pass is then output-stationary: one lane per
// output column adds that column's products in order (separate multiply and add, as everywhere) - no hash
// table, no atomics, no shared-memory read-modify-write chain, no synchronisation inside a row.
//
// Rows whose kept set left K+ (or that have no structure yet) are appended to a list that the row-group kernel
// of the same step walks; it rebuilds their structures (build_row_structures below).
#pragma once
#include "propagate.cuh"

namespace crwr {

constexpr int kFastWarps = 4;              // warps per CTA of the fast kernel

// Pool block of one row: [kcols nkp x i32][kmask ceil(nkp/32) x u32][ocols no x i32][cptr (no+1) x i32]
//                        [okidx no x u16][ckidx nprod x u16](pad to 8)[cw nprod x T]
template <typename T>
struct StructView {
    const int32_t* kcols;
    const uint32_t* kmask;
    const int32_t* ocols;
    const int32_t* cptr;
    const uint16_t* okidx;
    const uint16_t* ckidx;
    const T* cw;
};
__host__ __device__ inline size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }
template <typename T>
__host__ __device__ inline size_t struct_bytes(int nkp, int no, int nprod) {
    size_t b = 4 * (size_t)nkp + 4 * (size_t)((nkp + 31) >> 5) + 4 * (size_t)no + 4 * (size_t)(no + 1) + 2 * (size_t)no + 2 * (size_t)nprod;
    return align8(b) + align8(sizeof(T) * (size_t)nprod);
}
template <typename T>
__device__ __forceinline__ StructView<T> struct_view(const unsigned char* p, int nkp, int no, int nprod) {
    StructView<T> v;
    v.kcols = reinterpret_cast<const int32_t*>(p);   p += 4 * (size_t)nkp;
    v.kmask = reinterpret_cast<const uint32_t*>(p);  p += 4 * (size_t)((nkp + 31) >> 5);
    v.ocols = reinterpret_cast<const int32_t*>(p);   p += 4 * (size_t)no;
    v.cptr = reinterpret_cast<const int32_t*>(p);    p += 4 * (size_t)(no + 1);
    v.okidx = reinterpret_cast<const uint16_t*>(p);  p += 2 * (size_t)no;
    v.ckidx = reinterpret_cast<const uint16_t*>(p);  p += 2 * (size_t)nprod;
    const size_t head = 4 * (size_t)nkp + 4 * (size_t)((nkp + 31) >> 5) + 4 * (size_t)no + 4 * (size_t)(no + 1) + 2 * (size_t)no + 2 * (size_t)nprod;
    p += align8(head) - head;
    v.cw = reinterpret_cast<const T*>(p);
    return v;
}

// ---------------------------------------------------------------------------------------------
// The fast kernel.  CTA b owns the R = kFastWarps * rpw consecutive rows starting at b * R (static schedule):
//   A  validate: header present and current, then every kept entry of the operand row must be a member of K+
//      (binary search in the shared-memory copy of K+); q[] receives the kept values, 0 elsewhere
//   -  one reservation of the product buffer and one append to the slow list per CTA
//   B  per valid row: out[s] = sum of its contributions in order, scale, restart, norm terms, selection
//      bookkeeping (histogram or window), compacted write
// ---------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ inline size_t fast_smem_bytes(int km, int rpw) {
    const size_t R = (size_t)kFastWarps * rpw;
    return align16(R * km * 4) + align16(R * km * sizeof(T)) + align16(R * sizeof(RowStruct)) + align16(R * 8) + 2 * align16(R * 4) +
           align16((size_t)kHistWin * 4) + 64;
}

template <typename T>
__global__ void __launch_bounds__(kFastWarps * 32) propagate_fast_kernel(PropArgs<T> a, int KM, int RPW) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int R = kFastWarps * RPW;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    unsigned char* p = smem_raw;
    int32_t* s_kc = reinterpret_cast<int32_t*>(p);       p += align16((size_t)R * KM * 4);
    T* s_qk = reinterpret_cast<T*>(p);                   p += align16((size_t)R * KM * sizeof(T));
    RowStruct* s_hdr = reinterpret_cast<RowStruct*>(p);  p += align16((size_t)R * sizeof(RowStruct));
    long long* s_base = reinterpret_cast<long long*>(p); p += align16((size_t)R * 8);
    int* s_ok = reinterpret_cast<int*>(p);               p += align16((size_t)R * 4);
    int* s_nk = reinterpret_cast<int*>(p);               p += align16((size_t)R * 4);
    unsigned int* win = reinterpret_cast<unsigned int*>(p); p += align16((size_t)kHistWin * 4);
    unsigned int* s_wn = reinterpret_cast<unsigned int*>(p);
    int* s_fits = reinterpret_cast<int*>(p + 4);

    for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) win[i] = 0u;
    if (threadIdx.x == 0) { *s_wn = 0u; *s_fits = 1; }
    pdl_wait();
    WalkState* st = a.st;
    if (st->done) return;
    const bool cut_on = __ldcg(&st->cut_on) != 0;
    const T cut = (T)__ldcg(&st->cut);
    const unsigned gen = __ldcg(&st->pool_gen);
    const bool win_mode = a.win_mode && __ldcg(&st->win_on) != 0;
    const unsigned long long wlo = __ldcg(&st->win_lo), whi = __ldcg(&st->win_hi);
    const bool pred_valid = a.prefill_l1 && __ldcg(&st->pred_valid) != 0;
    const typename KeyOf<T>::type pred0 = (typename KeyOf<T>::type)__ldcg(&st->pred_prefix0);
    const WinStage stage = {reinterpret_cast<unsigned long long*>(win), s_wn};
    WinAcc wacc = {0ull, ~0ull};

    const int row0 = blockIdx.x * R;
    const int nrows = min(R, a.n_rows - row0);

    // ---- A0: headers
    if ((int)threadIdx.x < R) {
        const int r = threadIdx.x;
        RowStruct h = {0, 0, 0, 0, 0u};
        if (r < nrows) h = a.rs[row0 + r];
        s_hdr[r] = h;
        s_ok[r] = (r < nrows && cut_on && h.gen == gen && gen != 0u && h.nkp <= KM) ? 1 : 0;
        s_nk[r] = 0;
    }
    __syncthreads();

    // ---- A1: membership of the kept entries in K+, q values
    for (int i = 0; i < RPW; ++i) {
        const int r = warp * RPW + i;
        if (r >= nrows || !s_ok[r]) continue;
        const RowStruct h = s_hdr[r];
        const int row = row0 + r;
        const int32_t* kcols = reinterpret_cast<const int32_t*>(a.pool + h.off);
        int32_t* kc = s_kc + (size_t)r * KM;
        T* qk = s_qk + (size_t)r * KM;
        for (int k = lane; k < h.nkp; k += 32) { kc[k] = __ldg(kcols + k); qk[k] = (T)0; }
        __syncwarp();
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        const int32_t* rc = a.in.cols + ptr;
        const T* rv = a.in.vals + ptr;
        bool bad = false;
        int nkept = 0;
        for (int b = 0; b < len; b += 32) {
            const int e = b + lane;
            bool keep = false;
            T v = (T)0;
            if (e < len) { v = __ldg(rv + e); keep = v > cut; }
            if (keep) {
                const int32_t c = __ldg(rc + e);
                int lo = 0, hi = h.nkp;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (kc[mid] < c) lo = mid + 1; else hi = mid;
                }
                if (lo < h.nkp && kc[lo] == c) qk[lo] = v; else bad = true;
            }
            nkept += __popc(__ballot_sync(kFull, keep));
        }
        bad = __any_sync(kFull, bad);
        if (lane == 0) { s_ok[r] = bad ? 0 : 1; s_nk[r] = nkept; }
    }
    __syncthreads();

    // ---- one reservation of the product buffer, one append to the slow list
    if (warp == 0) {
        const int r = lane;
        const bool ok = r < nrows && s_ok[r] != 0;
        const bool inv = r < nrows && s_ok[r] == 0;
        const int no = ok ? s_hdr[r].no : 0;
        int incl = no;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        long long base = 0;
        if (lane == 0 && total > 0) base = (long long)atomicAdd(&st->cursor, (unsigned long long)total);
        base = __shfl_sync(kFull, base, 0);
        if (r < R) s_base[r] = base + incl - no;
        if (lane == 0 && base + total > a.out_cap) { st->capacity_overflow = 1; *s_fits = 0; }
        const unsigned im = __ballot_sync(kFull, inv);
        const unsigned om = __ballot_sync(kFull, ok);
        if (im) {
            unsigned sb = 0;
            if (lane == 0) sb = atomicAdd(&st->slow_count, (unsigned)__popc(im));
            sb = __shfl_sync(kFull, sb, 0);
            if (inv) a.slow_list[sb + __popc(im & lanemask_lt())] = row0 + r;
        }
        if (lane == 0 && om) atomicAdd(&st->fast_rows, (unsigned)__popc(om));
    }
    __syncthreads();
    const bool fits = *s_fits != 0;

    // ---- B: numeric pass
    unsigned long long t_sq0 = 0, t_sq1 = 0, t_sq2 = 0, t_nnz = 0, t_kept = 0;
    int t_max_kept = 0, t_max_len = 0;
    for (int i = 0; i < RPW; ++i) {
        const int r = warp * RPW + i;
        if (r >= nrows || !s_ok[r]) continue;
        const RowStruct h = s_hdr[r];
        const int row = row0 + r;
        const StructView<T> sv = struct_view<T>(a.pool + h.off, h.nkp, h.no, h.nprod);
        const T* qk = s_qk + (size_t)r * KM;
        const int32_t dcol = a.row_begin + row;
        const long long base = s_base[r];
        Sq128 pos = {0ull, 0ull};
        int written = 0;
        for (int b = 0; b < h.no; b += 32) {
            const int s = b + lane;
            const bool act = s < h.no;
            T v = (T)0;
            int32_t col = -1;
            if (act) {
                const int cb = __ldg(sv.cptr + s), ce = __ldg(sv.cptr + s + 1);
                T acc = (T)0;
                for (int c = cb; c < ce; ++c) acc = add_rn(acc, mul_rn(qk[__ldg(sv.ckidx + c)], __ldg(sv.cw + c)));
                v = mul_rn(a.scale, acc);
                col = __ldg(sv.ocols + s);
                if (col == dcol) v = add_rn(v, a.restart);
                const unsigned ok_ = __ldg(sv.okidx + s);
                const T old = ok_ != 0xFFFFu ? qk[ok_] : (T)0;
                if (old != (T)0) { const T d = sub_rn(old, v); sq_add(pos, (double)d * (double)d); }
                else sq_add(pos, (double)v * (double)v);
            }
            const bool nzv = act && v != (T)0;
            if (win_mode) win_classify<T>(nzv, v, wlo, whi, wacc, stage, a.cand, a.cand_cap);
            else if (a.hist) hist_values<T>(a.hist, win, nzv, v, pred_valid, pred0);
            const unsigned m = __ballot_sync(kFull, nzv);
            if (nzv && fits) {
                const long long o = base + written + __popc(m & lanemask_lt());
                CRWR_CHECK(st, o >= 0 && o < a.out_cap, 6);
                a.out.cols[o] = col;
                a.out.vals[o] = v;
            }
            written += __popc(m);
        }
        // kept operand columns that are not output columns: their new value is 0
        for (int k = lane; k < h.nkp; k += 32) {
            const T q = qk[k];
            if (q != (T)0 && ((__ldg(sv.kmask + (k >> 5)) >> (k & 31)) & 1u) == 0u) sq_add(pos, (double)q * (double)q);
        }
        const Sq128 rowsq = sq_group_sum(pos, 32);
        if (fits) for (int k = written + lane; k < h.no; k += 32) { a.out.cols[base + k] = -1; a.out.vals[base + k] = (T)0; }
        if (lane == 0) {
            a.out.row_ptr[row] = fits ? base : 0;
            a.out.row_len[row] = fits ? written : 0;
            t_sq2 += rowsq.hi;
            t_sq1 += rowsq.lo >> 32;
            t_sq0 += rowsq.lo & 0xffffffffull;
            t_kept += (unsigned long long)s_nk[r];
            t_nnz += (unsigned long long)written;
            t_max_len = max(t_max_len, written);
            t_max_kept = max(t_max_kept, h.nkp);
        }
    }
    pdl_launch_dependents();

    // ---- publish the CTA's totals (lane 0 of every warp holds its rows' sums)
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
        if (s_tot[0]) atomicAdd(&st->sq_limb[0], s_tot[0]);
        if (s_tot[1]) atomicAdd(&st->sq_limb[1], s_tot[1]);
        if (s_tot[2]) atomicAdd(&st->sq_limb[2], s_tot[2]);
        if (s_tot[3]) atomicAdd(&st->nnz_exact, s_tot[3]);
        if (s_tot[4]) atomicAdd(&st->nnz_in, s_tot[4]);
        if (s_max[0]) atomicMax(&st->max_kept, s_max[0]);
        if (s_max[1]) atomicMax(&st->max_row_len, s_max[1]);
    }
    if (win_mode) win_flush(wacc, stage, a.cand, a.cand_cap, st);
    else if (a.hist) hist_window_flush<T>(a.hist, win);
}

}  // namespace crwr
