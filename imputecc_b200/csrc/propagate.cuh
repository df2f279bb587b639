// K2: one walk step's propagation  Q~ = (1-rp) Q P + rp I   (utility.py:282) fused with the
// lazy percentile cut of the previous step (utility.py:290, applied while the row is read),
// the restart add, the zero drop of csr_matmat/csr_plus_csr and the per-row contribution to
// ||Q - Q~||_F (utility.py:283).
//
// Work decomposition: one warp owns one restart row (rows are independent; they are handed
// out by an atomic row counter).  The warp walks the kept entries k of its row IN ORDER and,
// for each, its 32 lanes stream row k of P (coalesced; distinct columns, so no two lanes meet
// in one accumulator slot).  Accumulators live in a per-warp open-addressing hash table in
// shared memory; a first-touch list records the order in which columns appeared.
//
// Why "in order": scipy's Gustavson kernel adds the products of one output entry in the
// storage order of the left operand's row, with a separate multiply and add.  Keeping that
// order (sorted columns after a cut - `Q > cut` sorts the operand in place, scipy
// _compressed.py:_scalar_binopt - storage order otherwise) makes every stored value
// bit-identical to the reference's, in float32 and in float64, and independent of scheduling
// and of how rows are sharded over GPUs.  The output row is written as scipy would store it:
// [restart entry if it is a new column] followed by the columns in first-touch order.
#pragma once
#include "common.cuh"

namespace crwr {

template <typename T>
struct PropArgs {
    const int32_t* p_indptr;
    const int32_t* p_cols;
    const T* p_vals;
    IterBuf<T> in, out;
    long long out_cap;
    WalkState* st;
    int32_t* fb_rows;        // rows deferred to the large-row kernel
    int32_t n_rows;          // rows of this shard
    int32_t row_begin;       // global index of local row 0 (= its restart column)
    T scale;                 // (T)(1 - rp)
    T restart;               // (T)(float)rp : rp * I is float32 in the reference (utility.py:278)
};

// Adds one row's share of ||Q - Q~||_F^2 to the step total as three 32-bit limbs of a 2^-64
// fixed-point number.  Integer sums commute, so delta is bit-identical whatever the order in
// which rows finish and however the rows are sharded over GPUs.
__device__ __forceinline__ void add_rowsq(WalkState* st, double x) {
    if (!(x > 0.0)) return;
    if (x >= 4294967296.0) x = 4294967295.0;
    const double hi = floor(x);
    const double f1 = (x - hi) * 4294967296.0;
    const double m1 = floor(f1);
    const double m0 = floor((f1 - m1) * 4294967296.0);
    atomicAdd(&st->sq_limb[2], (unsigned long long)hi);
    atomicAdd(&st->sq_limb[1], (unsigned long long)m1);
    atomicAdd(&st->sq_limb[0], (unsigned long long)m0);
}

__device__ __forceinline__ uint32_t hash_col(int32_t j, int log_h) {
    return ((uint32_t)j * 0x9E3779B1u) >> (32 - log_h);
}

// Writes one finished row.  `emit(i, &col, &val)` yields the i-th candidate (i < n_cand) in output
// order; zero values are dropped.  Returns nothing; updates cursor / row tables / overflow flag.
template <typename T, typename Emit>
__device__ __forceinline__ void write_row(const PropArgs<T>& a, int row, int n_cand, int n_nonzero, Emit emit) {
    const int lane = lane_id();
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(&a.st->cursor, (unsigned long long)n_nonzero);
    base = __shfl_sync(kFull, base, 0);
    const bool fits = base + n_nonzero <= a.out_cap;
    if (lane == 0) {
        a.out.row_ptr[row] = fits ? base : 0;
        a.out.row_len[row] = fits ? n_nonzero : 0;
        if (!fits) a.st->capacity_overflow = 1;
        atomicMax(&a.st->max_row_len, n_nonzero);
    }
    if (!fits) return;
    int written = 0;
    for (int b = 0; b < n_cand; b += 32) {
        int i = b + lane;
        int32_t c = 0;
        T v = (T)0;
        if (i < n_cand) emit(i, c, v);
        bool nz = (i < n_cand) && (v != (T)0);
        unsigned m = __ballot_sync(kFull, nz);
        if (nz) {
            long long o = base + written + __popc(m & lanemask_lt());
            a.out.cols[o] = c;
            a.out.vals[o] = v;
        }
        written += __popc(m);
    }
}

// ---------------------------------------------------------------------------------------------
// Main kernel: shared-memory hash accumulator, sized at run time.
//
// Per-warp shared memory (PropShape): accumulator table of H slots (keys + values), first-touch
// list, the kept operand entries with the extent of their P rows, and a double-buffered staging
// area into which the P rows of the next batch of up to 32 operand entries are copied with
// cp.async while the current batch is being accumulated - the dependent chain of a row then runs
// at shared-memory latency, not at L2/HBM latency.  Rows exceeding the table or the kept list
// are deferred to the large-row kernel.
// ---------------------------------------------------------------------------------------------
struct PropShape {
    int H;        // accumulator slots
    int limit;    // max distinct output columns held (<= 3/4 H)
    int kmax;     // max kept operand entries
    int scap;     // staging capacity (products) per buffer
    int warps;    // warps per CTA
    size_t per_warp;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T>
__host__ __device__ inline size_t prop_smem_per_warp(int H, int limit, int kmax, int scap) {
    return align16((size_t)H * sizeof(T)) + align16((size_t)kmax * sizeof(T)) + align16((size_t)2 * scap * sizeof(T)) +
           align16((size_t)H * 4) + 3 * align16((size_t)kmax * 4) + align16((size_t)2 * scap * 4) +
           align16((size_t)(limit + 32) * 2);
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_val(float* d, const float* g) { cp_async4(d, g); }
__device__ __forceinline__ void cp_async_val(double* d, const double* g) { cp_async8(d, g); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t hash_slot(int32_t j, int H) {
    return __umulhi((uint32_t)j * 0x9E3779B1u, (uint32_t)H);
}

// One batch of operand entries whose P rows are staged together.
struct Batch {
    int nb;      // operand entries in the batch (0: none)
    int len;     // this lane's entry: length of its P row   (lane t <-> entry first+t)
    int off;     // this lane's entry: offset of its products inside the staging buffer
    int direct;  // 1: a single entry whose row exceeds the staging buffer; streamed from global
};

template <typename T>
__global__ void __launch_bounds__(256) propagate_kernel(PropArgs<T> a, PropShape sh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (a.st->done) return;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int H = sh.H, KMAX = sh.kmax, SCAP = sh.scap, LIMIT = sh.limit;
    unsigned char* p = smem_raw + (size_t)warp * sh.per_warp;
    T* tval = reinterpret_cast<T*>(p);            p += align16((size_t)H * sizeof(T));
    T* kval = reinterpret_cast<T*>(p);            p += align16((size_t)KMAX * sizeof(T));
    T* sval = reinterpret_cast<T*>(p);            p += align16((size_t)2 * SCAP * sizeof(T));
    int32_t* tkey = reinterpret_cast<int32_t*>(p); p += align16((size_t)H * 4);
    int32_t* kcol = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* kps = reinterpret_cast<int32_t*>(p);  p += align16((size_t)KMAX * 4);
    int32_t* klen = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* scol = reinterpret_cast<int32_t*>(p); p += align16((size_t)2 * SCAP * 4);
    uint16_t* ord = reinterpret_cast<uint16_t*>(p);

    for (int i = lane; i < H; i += 32) tkey[i] = -1;
    __syncwarp();

    const bool cut_on = a.st->cut_on != 0;
    const T cut = (T)a.st->cut;
    unsigned long long kept_total = 0;

    // Picks the batch starting at operand entry `first` and issues the cp.async copies of its P
    // rows into staging buffer `buf`.  Uniform across the warp.
    auto stage = [&](int first, int nk, int buf) -> Batch {
        Batch bt;
        bt.nb = 0; bt.len = 0; bt.off = 0; bt.direct = 0;
        if (first >= nk) return bt;
        const int idx = first + lane;
        const int len = idx < nk ? klen[idx] : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned fit = __ballot_sync(kFull, idx < nk && incl <= SCAP);
        // entries are taken in order: the batch is the longest prefix that fits
        int nb = __ffs(~fit) - 1;                 // number of leading set bits
        if (fit == kFull) nb = 32;
        bt.len = len;
        bt.off = incl - len;
        if (nb == 0) { bt.nb = 1; bt.direct = 1; return bt; }
        bt.nb = nb;
        const int ps = idx < nk ? kps[idx] : 0;
        int32_t* dc = scol + buf * SCAP;
        T* dv = sval + buf * SCAP;
        for (int t = 0; t < nb; ++t) {
            const int s_ = __shfl_sync(kFull, ps, t);
            const int l_ = __shfl_sync(kFull, len, t);
            const int o_ = __shfl_sync(kFull, bt.off, t);
            for (int e = lane; e < l_; e += 32) {
                cp_async4(dc + o_ + e, a.p_cols + s_ + e);
                cp_async_val(dv + o_ + e, a.p_vals + s_ + e);
            }
        }
        return bt;
    };

    for (;;) {
        int row = 0;
        if (lane == 0) row = (int)atomicAdd(&a.st->row_counter, 1u);
        row = __shfl_sync(kFull, row, 0);
        if (row >= a.n_rows) break;

        // ---- read the row, dropping entries at or below the previous cut (strict >, utility.py:290)
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        int nk = 0;
        for (int b = 0; b < len; b += 128) {
            int32_t c[4];
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * 32 + lane;
                c[u] = 0; v[u] = (T)0;
                if (e < len) { c[u] = a.in.cols[ptr + e]; v[u] = a.in.vals[ptr + e]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * 32 + lane;
                const bool keep = (e < len) && (!cut_on || v[u] > cut);
                const unsigned m = __ballot_sync(kFull, keep);
                const int pos = nk + __popc(m & lanemask_lt());
                if (keep && pos < KMAX) { kcol[pos] = c[u]; kval[pos] = v[u]; }
                nk += __popc(m);
            }
        }
        __syncwarp();
        if (lane == 0) atomicMax(&a.st->max_kept, nk);
        bool defer = nk > KMAX;
        int count = 0;
        if (!defer) {
            if (cut_on) warp_sort_pairs(kcol, kval, nk);
            for (int idx = lane; idx < nk; idx += 32) {
                const int k = kcol[idx];
                const int ps = __ldg(a.p_indptr + k);
                kps[idx] = ps;
                klen[idx] = __ldg(a.p_indptr + k + 1) - ps;
            }
            __syncwarp();

            // ---- accumulate products in order of k, P rows arriving through the staging buffers
            int first = 0, buf = 0;
            Batch cur = stage(first, nk, buf);
            cp_async_commit();
            while (cur.nb > 0 && !defer) {
                const int next_first = first + cur.nb;
                Batch nxt = stage(next_first, nk, buf ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const T myq = (first + lane < nk) ? kval[first + lane] : (T)0;
                const int32_t* bc = scol + buf * SCAP;
                const T* bv = sval + buf * SCAP;
                for (int t = 0; t < cur.nb && !defer; ++t) {
                    const T q = __shfl_sync(kFull, myq, t);
                    const int l_ = __shfl_sync(kFull, cur.len, t);
                    const int o_ = __shfl_sync(kFull, cur.off, t);
                    const int gs = cur.direct ? kps[first] : 0;
                    for (int e0 = 0; e0 < l_; e0 += 32) {
                        if (count + 32 > LIMIT) { defer = true; break; }
                        const int e = e0 + lane;
                        bool is_new = false;
                        uint32_t slot = 0;
                        if (e < l_) {
                            int32_t j;
                            T pv;
                            if (cur.direct) { j = __ldg(a.p_cols + gs + e); pv = __ldg(a.p_vals + gs + e); }
                            else { j = bc[o_ + e]; pv = bv[o_ + e]; }
                            const T prod = mul_rn(q, pv);
                            slot = hash_slot(j, H);
                            for (;;) {
                                const int32_t kk = tkey[slot];
                                const T curv = tval[slot];
                                if (kk == j) { tval[slot] = add_rn(curv, prod); break; }
                                if (kk == -1) {
                                    const int32_t old = atomicCAS(&tkey[slot], -1, j);
                                    if (old == -1) { tval[slot] = prod; is_new = true; break; }
                                }
                                slot = slot + 1 == (uint32_t)H ? 0u : slot + 1;
                            }
                        }
                        const unsigned nm = __ballot_sync(kFull, is_new);
                        if (is_new) ord[count + __popc(nm & lanemask_lt())] = (uint16_t)slot;
                        count += __popc(nm);
                        __syncwarp();
                    }
                }
                first = next_first;
                buf ^= 1;
                cur = nxt;
            }
            cp_async_wait<0>();
            __syncwarp();
        }
        if (defer) {
            // undo and hand the row to the large-row kernel
            for (int i = lane; i < count; i += 32) tkey[ord[i]] = -1;
            __syncwarp();
            if (lane == 0) a.fb_rows[atomicAdd(&a.st->fb_count, 1u)] = row;
            continue;
        }
        kept_total += (unsigned long long)nk;

        // ---- epilogue: scale, restart add, norm, write
        const int32_t dcol = a.row_begin + row;
        int dslot = -1;
        {
            uint32_t slot = hash_slot(dcol, H);
            for (;;) {
                const int32_t kk = tkey[slot];
                if (kk == dcol) { dslot = (int)slot; break; }
                if (kk == -1) break;
                slot = slot + 1 == (uint32_t)H ? 0u : slot + 1;
            }
        }
        const bool diag_new = dslot < 0;
        double s1 = 0.0;
        int nz = 0;
        for (int b = 0; b < count; b += 32) {
            const int i = b + lane;
            bool one = false;
            if (i < count) {
                const int slot = ord[i];
                T v = mul_rn(a.scale, tval[slot]);
                if (slot == dslot) v = add_rn(v, a.restart);
                tval[slot] = v;
                s1 += (double)v * (double)v;
                one = v != (T)0;
            }
            nz += __popc(__ballot_sync(kFull, one));
        }
        if (diag_new) {
            if (lane == 0) s1 += (double)a.restart * (double)a.restart;
            if (a.restart != (T)0) nz += 1;
        }
        __syncwarp();
        double s2 = 0.0;
        for (int idx = lane; idx < nk; idx += 32) {
            const int32_t j = kcol[idx];
            const T oldv = kval[idx];
            uint32_t slot = hash_slot(j, H);
            bool found = false;
            for (;;) {
                const int32_t kk = tkey[slot];
                if (kk == j) { found = true; break; }
                if (kk == -1) break;
                slot = slot + 1 == (uint32_t)H ? 0u : slot + 1;
            }
            T newv = (T)0;
            if (found) newv = tval[slot];
            else if (j == dcol) newv = a.restart;
            const T d = sub_rn(oldv, newv);
            s2 += (double)d * (double)d - (double)newv * (double)newv;
        }
        const double tot = warp_sum(s1 + s2);
        if (lane == 0) add_rowsq(a.st, tot);

        const int n_cand = count + (diag_new ? 1 : 0);
        const int shift = diag_new ? 1 : 0;
        write_row(a, row, n_cand, nz, [&](int i, int32_t& c, T& v) {
            if (i < shift) { c = dcol; v = a.restart; }
            else { const int slot = ord[i - shift]; c = tkey[slot]; v = tval[slot]; }
        });
        __syncwarp();
        for (int i = lane; i < count; i += 32) tkey[ord[i]] = -1;
        __syncwarp();
    }
    if (lane == 0 && kept_total) atomicAdd(&a.st->nnz_in, kept_total);
}

// ---------------------------------------------------------------------------------------------
// Large-row kernel: same algorithm with a dense accumulator of N entries per warp in global
// memory (L2 resident), for rows whose kept list or output does not fit the hash variant.
// Sorted traversal of the kept entries comes from a bitmap scan instead of a sort.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DenseScratch {
    T* acc;            // [warps][N]
    T* oldv;           // [warps][N]
    int32_t* ord;      // [warps][N]
    uint32_t* tbits;   // [warps][words]   touched columns
    uint32_t* kbits;   // [warps][words]   kept columns of the input row
    long long n;       // N
    long long words;   // ceil(N/32)
};

template <typename T>
__device__ __forceinline__ void dense_accumulate_row_of_p(const PropArgs<T>& a, T q, int k, T* acc, uint32_t* tbits,
                                                          int32_t* ord, int& count) {
    const int lane = lane_id();
    const int s = __ldg(a.p_indptr + k);
    const int e_end = __ldg(a.p_indptr + k + 1);
    for (int e0 = s; e0 < e_end; e0 += 32) {
        const int e = e0 + lane;
        bool is_new = false;
        int32_t j = 0;
        if (e < e_end) {
            j = __ldg(a.p_cols + e);
            const T prod = mul_rn(q, __ldg(a.p_vals + e));
            const uint32_t bit = 1u << (j & 31);
            const uint32_t old = atomicOr(&tbits[j >> 5], bit);
            if (old & bit) acc[j] = add_rn(acc[j], prod);
            else { acc[j] = prod; is_new = true; }
        }
        unsigned nm = __ballot_sync(kFull, is_new);
        if (is_new) ord[count + __popc(nm & lanemask_lt())] = j;
        count += __popc(nm);
        __syncwarp();
    }
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) propagate_dense_kernel(PropArgs<T> a, DenseScratch<T> sc) {
    if (a.st->done) return;
    const unsigned n_fb = a.st->fb_count;
    if (n_fb == 0) return;
    const int lane = lane_id();
    const long long gw = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
    T* acc = sc.acc + gw * sc.n;
    T* oldv = sc.oldv + gw * sc.n;
    int32_t* ord = sc.ord + gw * sc.n;
    uint32_t* tbits = sc.tbits + gw * sc.words;
    uint32_t* kbits = sc.kbits + gw * sc.words;
    const bool cut_on = a.st->cut_on != 0;
    const T cut = (T)a.st->cut;
    unsigned long long kept_total = 0;

    for (;;) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&a.st->fb_counter, 1u);
        r = __shfl_sync(kFull, r, 0);
        if (r >= n_fb) break;
        const int row = a.fb_rows[r];
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        int count = 0;
        int nk = 0;
        if (!cut_on) {
            // no cut on the operand: scipy consumes it in storage order
            nk = len;
            for (int e = 0; e < len; ++e) {
                const int k = a.in.cols[ptr + e];
                const T q = a.in.vals[ptr + e];
                if (lane == 0) { oldv[k] = q; atomicOr(&kbits[k >> 5], 1u << (k & 31)); }
                dense_accumulate_row_of_p(a, q, k, acc, tbits, ord, count);
            }
        } else {
            for (int b = 0; b < len; b += 32) {
                int e = b + lane;
                bool keep = false;
                if (e < len) {
                    const T v = a.in.vals[ptr + e];
                    if (v > cut) {
                        const int c = a.in.cols[ptr + e];
                        oldv[c] = v;
                        atomicOr(&kbits[c >> 5], 1u << (c & 31));
                        keep = true;
                    }
                }
                nk += __popc(__ballot_sync(kFull, keep));
            }
            __syncwarp();
            for (long long w0 = 0; w0 < sc.words; w0 += 32) {
                const long long w = w0 + lane;
                uint32_t word = (w < sc.words) ? kbits[w] : 0u;
                unsigned nzmask = __ballot_sync(kFull, word != 0u);
                while (nzmask) {
                    const int l = __ffs(nzmask) - 1;
                    nzmask &= nzmask - 1;
                    uint32_t bits = __shfl_sync(kFull, word, l);
                    while (bits) {
                        const int bpos = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int k = (int)((w0 + l) * 32 + bpos);
                        const T q = oldv[k];
                        dense_accumulate_row_of_p(a, q, k, acc, tbits, ord, count);
                    }
                }
            }
        }
        kept_total += (unsigned long long)nk;
        if (lane == 0) atomicMax(&a.st->max_kept, nk);
        __syncwarp();

        const int32_t dcol = a.row_begin + row;
        const bool diag_new = ((tbits[dcol >> 5] >> (dcol & 31)) & 1u) == 0u;
        double s1 = 0.0;
        int nz = 0;
        for (int b = 0; b < count; b += 32) {
            int i = b + lane;
            bool one = false;
            if (i < count) {
                const int j = ord[i];
                T v = mul_rn(a.scale, acc[j]);
                if (j == dcol) v = add_rn(v, a.restart);
                acc[j] = v;
                s1 += (double)v * (double)v;
                one = v != (T)0;
            }
            nz += __popc(__ballot_sync(kFull, one));
        }
        if (diag_new) {
            if (lane == 0) s1 += (double)a.restart * (double)a.restart;
            if (a.restart != (T)0) nz += 1;
        }
        __syncwarp();
        // old entries: every kept column has its bit in kbits and its value in oldv
        double s2 = 0.0;
        for (int b = 0; b < len; b += 32) {
            int e = b + lane;
            if (e < len) {
                const int j = a.in.cols[ptr + e];
                if ((kbits[j >> 5] >> (j & 31)) & 1u) {
                    const T o = oldv[j];
                    T newv = (T)0;
                    if ((tbits[j >> 5] >> (j & 31)) & 1u) newv = acc[j];
                    else if (j == dcol) newv = a.restart;
                    const T d = sub_rn(o, newv);
                    s2 += (double)d * (double)d - (double)newv * (double)newv;
                }
            }
        }
        const double tot = warp_sum(s1 + s2);
        if (lane == 0) add_rowsq(a.st, tot);

        const int n_cand = count + (diag_new ? 1 : 0);
        const int shift = diag_new ? 1 : 0;
        write_row(a, row, n_cand, nz, [&](int i, int32_t& c, T& v) {
            if (i < shift) { c = dcol; v = a.restart; }
            else { c = ord[i - shift]; v = acc[c]; }
        });
        __syncwarp();
        // clear the bitmaps through the lists that set them
        for (int i = lane; i < count; i += 32) tbits[ord[i] >> 5] = 0u;
        for (int b = lane; b < len; b += 32) kbits[a.in.cols[ptr + b] >> 5] = 0u;
        __syncwarp();
    }
    if (lane == 0 && kept_total) atomicAdd(&a.st->nnz_in, kept_total);
}

// Q0 = I (utility.py:279): every local row holds its restart column with value 1.
template <typename T>
__global__ void init_iterate_kernel(IterBuf<T> buf, int n_rows, int row_begin) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) {
        buf.cols[r] = row_begin + r;
        buf.vals[r] = (T)1;
        buf.row_ptr[r] = r;
        buf.row_len[r] = 1;
    }
}

}  // namespace crwr
