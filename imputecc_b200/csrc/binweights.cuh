// Bin-contact weights of match_contigs (utility.py:466-480): for every contig to bin and every existing bin,
//   hic_sum = 0;  for j in bins[b]: hic_sum += E[contig, bins[b][j]];  weight = -hic_sum / len(bins[b])
// with E the imputed m x m matrix.  The reference does |to_bin| x |binned contigs| scalar look-ups in a LIL matrix
// per iteration (its second CPU hot loop).
//
// One thread per (contig, bin): it walks the bin's members in list order and finds each in the contig's CSR row
// by binary search (rows of E are sorted), adding the hits one after the other - the float sum therefore runs in
// exactly the reference's order (a missing entry adds 0.0 there, which changes nothing), and the result carries the
// reference's sign of zero (-0.0 for a bin without contact).  No sort, no atomics, no shared state between threads.
#pragma once
#include "common.cuh"

namespace crwr {

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
    const long long s0 = __ldg(indptr + c), s1 = __ldg(indptr + c + 1);
    const long long m0 = __ldg(bin_ptr + b), m1 = __ldg(bin_ptr + b + 1);
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
    out[idx] = div_rn(-sum, (T)(m1 - m0));
}

}  // namespace crwr
