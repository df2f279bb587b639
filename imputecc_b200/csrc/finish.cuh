// Walk output: the m x m slice of the final iterate (utility.py:306-307) and the symmetric
// normalisation  E = Q + Q^T,  E = D^-1/2 E D^-1/2  (imputation.py:122-127).
#pragma once
#include "common.cuh"

namespace crwr {

// Entries of local row `row` that survive the last cut and lie in columns < col_limit.
template <typename T>
__global__ void result_count_kernel(IterBuf<T> buf, const WalkState* __restrict__ st, int n_rows, int col_limit,
                                    int32_t* __restrict__ counts) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int lane = lane_id();
    const bool cut_on = st->cut_on != 0;
    const T cut = (T)st->cut;
    const long long ptr = buf.row_ptr[row];
    const int len = buf.row_len[row];
    int cnt = 0;
    for (int e = lane; e < len; e += 32) {
        const T v = buf.vals[ptr + e];
        cnt += (buf.cols[ptr + e] < col_limit) && (!cut_on || v > cut);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if (lane == 0) counts[row] = cnt;
}

template <typename T>
__global__ void result_write_kernel(IterBuf<T> buf, const WalkState* __restrict__ st, int n_rows, int col_limit,
                                    const int64_t* __restrict__ indptr, int32_t* __restrict__ out_cols,
                                    T* __restrict__ out_vals) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    const int lane = lane_id();
    const bool cut_on = st->cut_on != 0;
    const T cut = (T)st->cut;
    const long long ptr = buf.row_ptr[row];
    const int len = buf.row_len[row];
    const long long base = indptr[row];
    int w = 0;
    for (int b = 0; b < len; b += 32) {
        const int e = b + lane;
        int c = 0;
        T v = (T)0;
        bool keep = false;
        if (e < len) {
            c = buf.cols[ptr + e];
            v = buf.vals[ptr + e];
            keep = (c < col_limit) && (!cut_on || v > cut);
        }
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) {
            const long long o = base + w + __popc(m & lanemask_lt());
            out_cols[o] = c;
            out_vals[o] = v;
        }
        w += __popc(m);
    }
    __syncwarp();
    if (w > 1) warp_sort_pairs_any(out_cols + base, out_vals + base, w);
}

// ---------------------------------------------------------------------------------------------
// Symmetrise + normalise
// ---------------------------------------------------------------------------------------------
template <typename T>
struct SymScratch {
    int32_t* t_cnt;      // [m+1]
    int64_t* t_ptr;      // [m+1]
    int32_t* t_fill;     // [m]
    int32_t* t_col;      // [nnz]  transposed structure: for column c, the rows holding it
    T* t_val;            // [nnz]
    int32_t* e_cnt;      // [m+1]
    int64_t* e_ptr;      // [m+1]
    T* b;                // [m]   1/sqrt(colsum)
    long long* total;    // [1]
};

template <typename T>
__global__ void sym_count_cols_kernel(int m, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                      SymScratch<T> sc) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= m) return;
    for (long long e = indptr[row] + lane_id(); e < indptr[row + 1]; e += 32) atomicAdd(&sc.t_cnt[indices[e]], 1);
}

template <typename T>
__global__ void sym_scatter_kernel(int m, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                   const T* __restrict__ data, SymScratch<T> sc) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= m) return;
    for (long long e = indptr[row] + lane_id(); e < indptr[row + 1]; e += 32) {
        const int c = indices[e];
        const long long pos = sc.t_ptr[c] + atomicAdd(&sc.t_fill[c], 1);
        sc.t_col[pos] = row;
        sc.t_val[pos] = data[e];
    }
}

__device__ __forceinline__ int lower_bound_i32(const int32_t* a, int n, int32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// sort each transposed row, then count |row_i(Q) U row_i(Q^T)|
template <typename T>
__global__ void sym_sort_count_kernel(int m, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                      SymScratch<T> sc) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= m) return;
    const int lane = lane_id();
    const long long ts = sc.t_ptr[row];
    const int tn = (int)(sc.t_ptr[row + 1] - ts);
    if (tn > 1) warp_sort_pairs_any(sc.t_col + ts, sc.t_val + ts, tn);
    __syncwarp();
    const long long as = indptr[row];
    const int an = (int)(indptr[row + 1] - as);
    int common = 0;
    for (int i = lane; i < an; i += 32) {
        const int c = indices[as + i];
        const int p = lower_bound_i32(sc.t_col + ts, tn, c);
        common += (p < tn && sc.t_col[ts + p] == c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) common += __shfl_xor_sync(kFull, common, o);
    if (lane == 0) sc.e_cnt[row] = an + tn - common;
}

// merged row i of E = Q + Q^T (sorted), its numpy-order sum, b_i = 1/sqrt(d_i)
template <typename T>
__global__ void sym_merge_kernel(int m, const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                 const T* __restrict__ data, SymScratch<T> sc, const int64_t* __restrict__ e_ptr,
                                 int32_t* __restrict__ out_cols, T* __restrict__ out_vals) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= m) return;
    const int lane = lane_id();
    const long long ts = sc.t_ptr[row];
    const int tn = (int)(sc.t_ptr[row + 1] - ts);
    const long long as = indptr[row];
    const int an = (int)(indptr[row + 1] - as);
    const long long base = e_ptr[row];
    const int en = (int)(e_ptr[row + 1] - base);
    // Q's entries, each plus its mirror when present  (E += E.T, imputation.py:123)
    for (int i = lane; i < an; i += 32) {
        const int c = indices[as + i];
        const int p = lower_bound_i32(sc.t_col + ts, tn, c);
        T v = data[as + i];
        if (p < tn && sc.t_col[ts + p] == c) v = add_rn(v, sc.t_val[ts + p]);
        out_cols[base + i] = c;
        out_vals[base + i] = v;
    }
    // mirror-only entries, appended, then the row is sorted by column
    int w = an;
    for (int k0 = 0; k0 < tn; k0 += 32) {
        const int k = k0 + lane;
        bool only = false;
        int c = 0;
        if (k < tn) {
            c = sc.t_col[ts + k];
            const int p = lower_bound_i32(indices + as, an, c);
            only = !(p < an && indices[as + p] == c);
        }
        const unsigned mk = __ballot_sync(kFull, only);
        if (only) {
            const long long o = base + w + __popc(mk & lanemask_lt());
            out_cols[o] = c;
            out_vals[o] = sc.t_val[ts + k];
        }
        w += __popc(mk);
    }
    __syncwarp();
    if (w > 1) warp_sort_pairs_any(out_cols + base, out_vals + base, w);
    __syncwarp();
    if (lane == 0) {
        T d = (T)0;
        if (en > 0) d = np_segment_sum<T>(out_vals + base, en);     // imputation.py:124
        if (d == (T)0) d = (T)1;                                    // :125
        sc.b[row] = div_rn((T)1, sqrt_rn(d));                       // :126
    }
}

template <typename T>
__global__ void sym_scale_kernel(int m, const int64_t* __restrict__ e_ptr, const int32_t* __restrict__ cols,
                                 T* __restrict__ vals, const T* __restrict__ b) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= m) return;
    const T bi = b[row];
    for (long long e = e_ptr[row] + lane_id(); e < e_ptr[row + 1]; e += 32)
        vals[e] = mul_rn(mul_rn(bi, vals[e]), b[cols[e]]);          // (b E) b, imputation.py:127
}

}  // namespace crwr
