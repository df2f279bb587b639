// Bin-contact weights of match_contigs (utility.py:466-480): for every contig to bin and every existing bin,
//   hic_sum = 0;  for j in bins[b]: hic_sum += E[contig, bins[b][j]];  weight = -hic_sum / len(bins[b])
// with E the imputed m x m matrix.  The reference does |to_bin| x |binned contigs| scalar look-ups in a LIL matrix
// per iteration (its second CPU hot loop).
//
// Exactness: the float sum of a (contig, bin) pair runs over the bin's members in list order; a member without a
// stored contact adds 0.0 in the reference, which changes nothing, so only the stored entries of the contig's row
// matter, taken in the order of their position in the bin's list.  A bin without contact gets -0.0 (= -(0.0)/n).
//
// Work proportional to the stored entries (bin_weights_rows_kernel, one warp per contig): the row's entries are
// tagged with (bin, position in the bin's list) through an inverse map built once per call (bin_inverse_kernel),
// sorted by that tag in shared memory, and every run of equal bin is summed front to back by one lane.  The output
// row is pre-filled with -0.0.  Rows longer than the shared-memory list, and callers without scratch for the inverse
// map, take the general form (bin_weights_kernel: one thread per (contig, bin) walks the bin's members in list order
// and finds each in the contig's sorted CSR row by binary search).
#pragma once
#include "common.cuh"

namespace crwr {

constexpr int kBinRowCap = 256;      // stored entries of a contig's row handled in shared memory
constexpr int kBinRowWarps = 4;

// One (contig, bin) pair by binary search over the contig's row, members in list order.
template <typename T>
__device__ __forceinline__ T bin_pair_weight(const int32_t* __restrict__ indices, const T* __restrict__ data, long long s0,
                                             long long s1, const int32_t* __restrict__ bin_members, long long m0, long long m1) {
    T sum = (T)0;
    for (long long j = m0; j < m1; ++j) {
        const int32_t col = __ldg(bin_members + j);
        long long lo = s0, hi = s1;                  // first position with indices[pos] >= col
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(indices + mid) < col) lo = mid + 1; else hi = mid;
        }
        if (lo < s1 && __ldg(indices + lo) == col) sum = add_rn(sum, __ldg(data + lo));
    }
    return div_rn(-sum, (T)(m1 - m0));
}

template <typename T>
__global__ void __launch_bounds__(256) bin_weights_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                          const T* __restrict__ data, const int32_t* __restrict__ rows,
                                                          long long n_rows, const int64_t* __restrict__ bin_ptr,
                                                          const int32_t* __restrict__ bin_members, long long n_bins,
                                                          T* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * n_bins) return;
    const long long r = idx / n_bins, b = idx - r * n_bins;
    const int c = __ldg(rows + r);
    out[idx] = bin_pair_weight<T>(indices, data, __ldg(indptr + c), __ldg(indptr + c + 1), bin_members, __ldg(bin_ptr + b),
                                  __ldg(bin_ptr + b + 1));
}

// bin_of[contig] / pos_of[contig] for every member of every bin (bins are disjoint); other contigs keep bin -1.
__global__ void bin_inverse_kernel(const int64_t* __restrict__ bin_ptr, const int32_t* __restrict__ bin_members, long long n_bins,
                                   int32_t* __restrict__ bin_of, int32_t* __restrict__ pos_of) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= bin_ptr[n_bins]) return;
    long long lo = 0, hi = n_bins;                   // last bin whose first member is at or before j
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (bin_ptr[mid] <= j) lo = mid; else hi = mid;
    }
    const int32_t c = bin_members[j];
    bin_of[c] = (int32_t)lo;
    pos_of[c] = (int32_t)(j - bin_ptr[lo]);
}

template <typename T>
__global__ void __launch_bounds__(kBinRowWarps * 32) bin_weights_rows_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const T* __restrict__ data,
    const int32_t* __restrict__ rows, long long n_rows, const int64_t* __restrict__ bin_ptr,
    const int32_t* __restrict__ bin_members, long long n_bins, const int32_t* __restrict__ bin_of,
    const int32_t* __restrict__ pos_of, T* __restrict__ out) {
    __shared__ unsigned long long s_key[kBinRowWarps][kBinRowCap];
    __shared__ T s_val[kBinRowWarps][kBinRowCap];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const long long r = (long long)blockIdx.x * kBinRowWarps + warp;
    if (r >= n_rows) return;
    const int c = __ldg(rows + r);
    const long long s0 = __ldg(indptr + c), s1 = __ldg(indptr + c + 1);
    T* orow = out + r * n_bins;
    if (s1 - s0 > kBinRowCap) {
        for (long long b = lane; b < n_bins; b += 32)
            orow[b] = bin_pair_weight<T>(indices, data, s0, s1, bin_members, __ldg(bin_ptr + b), __ldg(bin_ptr + b + 1));
        return;
    }
    for (long long b = lane; b < n_bins; b += 32) orow[b] = -(T)0;          // a bin without contact: -(0.0) / n
    unsigned long long* key = s_key[warp];
    T* val = s_val[warp];
    int n = 0;
    for (long long e0 = s0; e0 < s1; e0 += 32) {
        const long long e = e0 + lane;
        int b = -1;
        int32_t col = 0;
        if (e < s1) { col = __ldg(indices + e); b = __ldg(bin_of + col); }
        const bool keep = b >= 0;
        const unsigned mk = __ballot_sync(kFull, keep);
        if (keep) {
            const int p = n + __popc(mk & lanemask_lt());
            key[p] = ((unsigned long long)(unsigned)b << 32) | (unsigned)__ldg(pos_of + col);
            val[p] = __ldg(data + e);
        }
        n += __popc(mk);
    }
    __syncwarp();
    if (n > 1) warp_sort_pairs(key, val, n);                                // by (bin, position in the bin's list)
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const unsigned b = (unsigned)(key[i] >> 32);
        if (i > 0 && (unsigned)(key[i - 1] >> 32) == b) continue;           // not the first entry of its bin
        T sum = val[i];                                                     // 0 + x == x
        for (int j = i + 1; j < n && (unsigned)(key[j] >> 32) == b; ++j) sum = add_rn(sum, val[j]);
        orow[b] = div_rn(-sum, (T)(__ldg(bin_ptr + b + 1) - __ldg(bin_ptr + b)));
    }
}

}  // namespace crwr
