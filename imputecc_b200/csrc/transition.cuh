// K1: the walk matrix P of imputation.py:112-119, built on the device directly in the
// markers-first index space (the reference permutes physically, :112-113; here the relabelling
// new_of_old[] is applied while the rows are written).
//
//   A = M - diag(M)                       diagonal and explicit zeros are not stored
//   isolated_j = (colsum_j(A) == 0)       B = A + diag(isolated)
//   P[i,:] = (1 / colsum_i(B)) * B[i,:]   cast to the walk dtype
//
// Column sums follow numpy's pairwise order over the column's entries by ascending row
// (scipy sums a CSC matrix's columns with np.add.reduceat), so P is bit-identical to the
// reference's; that needs the transpose structure, built here by count / scan / scatter /
// per-column sort.
#pragma once
#include "common.cuh"

namespace crwr {

// Exclusive scan of int32 counts into OutT offsets (n+1 entries) by ONE CTA of 1024 threads, 8 items per
// thread and iteration (n is at most a few hundred thousand: one CTA keeps it a single short launch).
template <typename OutT>
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ cnt, OutT* __restrict__ out,
                                                              long long n, long long* total_out) {
    constexpr int ITEMS = 8;
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    const int t = threadIdx.x;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024 * ITEMS) {
        const long long i0 = base + (long long)t * ITEMS;
        long long v[ITEMS];
        long long mine = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            v[k] = (i0 + k < n) ? (long long)cnt[i0 + k] : 0;
            mine += v[k];
        }
        long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long w = __shfl_up_sync(kFull, incl, o);
            if ((t & 31) >= o) incl += w;
        }
        if ((t & 31) == 31) warp_tot[t >> 5] = incl;
        __syncthreads();
        if (t < 32) {
            long long w = warp_tot[t];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long x = __shfl_up_sync(kFull, wi, o);
                if (t >= o) wi += x;
            }
            warp_tot[t] = wi - w;
        }
        __syncthreads();
        long long run = carry_s + warp_tot[t >> 5] + incl - mine;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (i0 + k < n) out[i0 + k] = (OutT)run;
            run += v[k];
        }
        __syncthreads();
        if (t == 1023) carry_s = run;
        __syncthreads();
    }
    if (t == 0) {
        out[n] = (OutT)carry_s;
        if (total_out) *total_out = carry_s;
    }
}

// The same scan over many CTAs for long inputs (the single-CTA form takes ~40 us at 50,000 counts): pass 1 sums
// tiles of 8192 counts, pass 2 gives every tile the sum of the tiles before it and scans the tile.
constexpr int kScanTile = 8192;
constexpr int kScanMaxTiles = 4096;

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const int32_t* __restrict__ cnt, long long n,
                                                               long long* __restrict__ tile_sum) {
    __shared__ long long warp_tot[32];
    const int t = threadIdx.x;
    const long long i0 = (long long)blockIdx.x * kScanTile + (long long)t * 8;
    long long mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) mine += (i0 + k < n) ? (long long)cnt[i0 + k] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(kFull, mine, o);
    if ((t & 31) == 0) warp_tot[t >> 5] = mine;
    __syncthreads();
    if (t < 32) {
        long long w = warp_tot[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
        if (t == 0) tile_sum[blockIdx.x] = w;
    }
}

template <typename OutT>
__global__ void __launch_bounds__(1024) scan_tiles_kernel(const int32_t* __restrict__ cnt, OutT* __restrict__ out, long long n,
                                                          const long long* __restrict__ tile_sum, long long* total_out) {
    constexpr int ITEMS = 8;
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    const int t = threadIdx.x;
    // sum of the tiles before this one
    long long before = 0;
    for (int i = t; i < (int)blockIdx.x; i += 1024) before += tile_sum[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(kFull, before, o);
    if ((t & 31) == 0) warp_tot[t >> 5] = before;
    __syncthreads();
    if (t < 32) {
        long long w = warp_tot[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(kFull, w, o);
        if (t == 0) carry_s = w;
    }
    __syncthreads();
    const long long carry = carry_s;
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * kScanTile + (long long)t * ITEMS;
    long long v[ITEMS];
    long long mine = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        v[k] = (i0 + k < n) ? (long long)cnt[i0 + k] : 0;
        mine += v[k];
    }
    long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long w = __shfl_up_sync(kFull, incl, o);
        if ((t & 31) >= o) incl += w;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        long long w = warp_tot[t];
        long long wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long x = __shfl_up_sync(kFull, wi, o);
            if (t >= o) wi += x;
        }
        warp_tot[t] = wi - w;
    }
    __syncthreads();
    long long run = carry + warp_tot[t >> 5] + incl - mine;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (i0 + k < n) out[i0 + k] = (OutT)run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && t == 1023) {       // the last thread of the last tile holds the grand total
        out[n] = (OutT)run;
        if (total_out) *total_out = run;
    }
}

struct TransitionScratch {
    int32_t* row_cnt;     // [N+1] entries of B per new row
    int32_t* col_cnt;     // [N+1] entries of A per new column
    int32_t* col_ptr;     // [N+1]
    int32_t* col_fill;    // [N]
    int32_t* csc_row;     // [nnz]
    double* csc_val;      // [nnz]
    double* scale;        // [N]  1 / colsum(B)
    int32_t* isolated;    // [N]
    double* row_val;      // [nnz + N] unscaled float64 values of B in CSR order
    long long* total;     // [1]
    int32_t* bad;         // [1] set when a contact value is negative, NaN or infinite
};

__global__ void transition_count_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                        const double* __restrict__ data, const int32_t* __restrict__ new_of_old,
                                        long long n, TransitionScratch sc) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int cnt = 0;
    for (long long e = indptr[r]; e < indptr[r + 1]; ++e) {
        const int c = indices[e];
        const double v = data[e];
        if (!(v >= 0.0) || v > 1.7976931348623157e308) *sc.bad = 1;      // negative, NaN or +inf
        if (c != r && v != 0.0) {
            ++cnt;
            atomicAdd(&sc.col_cnt[new_of_old[c]], 1);
        }
    }
    sc.row_cnt[new_of_old[r]] = cnt;
}

__global__ void transition_scatter_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                          const double* __restrict__ data, const int32_t* __restrict__ new_of_old,
                                          long long n, TransitionScratch sc) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int rn = new_of_old[r];
    for (long long e = indptr[r]; e < indptr[r + 1]; ++e) {
        const int c = indices[e];
        const double v = data[e];
        if (c != r && v != 0.0) {
            const int cn = new_of_old[c];
            const int pos = sc.col_ptr[cn] + atomicAdd(&sc.col_fill[cn], 1);
            sc.csc_row[pos] = rn;
            sc.csc_val[pos] = v;
        }
    }
}

// One warp per (new) column: order its entries by row, sum them numpy-style, derive the scale.
__global__ void transition_colsum_kernel(long long n, TransitionScratch sc) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= n) return;
    const int s = sc.col_ptr[col];
    const int len = sc.col_ptr[col + 1] - s;
    if (len > 1) warp_sort_pairs_any(sc.csc_row + s, sc.csc_val + s, len);
    __syncwarp();
    if (lane_id() == 0) {
        double sum = 0.0;
        if (len > 0) sum = np_segment_sum<double>(sc.csc_val + s, len);
        const int iso = (sum == 0.0) ? 1 : 0;                  // imputation.py:117
        sc.isolated[col] = iso;
        // colsum(B): A's column plus the unit self-loop of an isolated node
        const double colsum_b = iso ? (len > 0 ? sum + 1.0 : 1.0) : sum;
        sc.scale[col] = div_rn(1.0, colsum_b);                 // imputation.py:118
        sc.row_cnt[col] += iso;                                // row `col` of B gains the self-loop
    }
}

// One warp per ORIGINAL row: relabel, drop diagonal/zeros, add the self-loop, sort by new column.
template <typename T>
__global__ void transition_fill_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const double* __restrict__ data, const int32_t* __restrict__ new_of_old,
                                       long long n, TransitionScratch sc, const int32_t* __restrict__ p_indptr,
                                       int32_t* __restrict__ p_cols, T* __restrict__ p_vals) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const int lane = lane_id();
    const int rn = new_of_old[r];
    const int base = p_indptr[rn];
    const long long s = indptr[r];
    const int len = (int)(indptr[r + 1] - s);
    int w = 0;
    for (int b = 0; b < len; b += 32) {
        const int e = b + lane;
        int c = 0;
        double v = 0.0;
        bool keep = false;
        if (e < len) {
            c = indices[s + e];
            v = data[s + e];
            keep = (c != r) && (v != 0.0);
        }
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) {
            const int o = base + w + __popc(m & lanemask_lt());
            p_cols[o] = new_of_old[c];
            sc.row_val[o] = v;
        }
        w += __popc(m);
    }
    if (sc.isolated[rn]) {
        if (lane == 0) { p_cols[base + w] = rn; sc.row_val[base + w] = 1.0; }
        w += 1;
    }
    __syncwarp();
    if (w > 1) warp_sort_pairs_any(p_cols + base, sc.row_val + base, w);
    __syncwarp();
    const double scl = sc.scale[rn];
    for (int i = lane; i < w; i += 32) p_vals[base + i] = (T)mul_rn(scl, sc.row_val[base + i]);   // :118-119
}

}  // namespace crwr
