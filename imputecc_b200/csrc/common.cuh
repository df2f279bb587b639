// Shared device-side definitions for libcrwr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/crwr_b200.h"

namespace crwr {

constexpr int kSelectBits = 12;                 // radix digit of the percentile select
constexpr int kSelectBins = 1 << kSelectBits;   // 4096
constexpr int kSelectLevels = 2;                // full-data histogram levels before the gather
constexpr int kCandSmem = 4096;                 // candidates ranked inside shared memory
constexpr int kMaxTrace = 4096;                 // per-step records kept on the device
constexpr unsigned kFull = 0xffffffffu;
constexpr int kHistWin = 512;                   // level-0 bins kept in a per-CTA shared-memory window

// ---------------------------------------------------------------------------------------------
// Exact (non-contracted) arithmetic.  scipy's csr_matmat does `sums[j] += v * B[k,j]` as a
// separate multiply and add (x86-64 baseline wheels carry no FMA), so these must never fuse.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// Order-preserving map float -> unsigned (all finite values, either sign).
// First level-0 bin of the per-CTA histogram window: 448 bins below bin(1.0), 64 at and above it
// (float32: 2^-56 .. 2^8 at 8 bins per octave; float64: 2^-448 .. 2^64 at one bin per octave).
template <typename T> struct HistWindow;
template <> struct HistWindow<float> { static constexpr int base = (0xBF800000u >> 20) - 448; };
template <> struct HistWindow<double> { static constexpr int base = 0xBFF - 448; };

template <typename T> struct KeyOf;
template <> struct KeyOf<float> {
    typedef uint32_t type;
    static constexpr int bits = 32;
    __device__ __forceinline__ static uint32_t enc(float v) {
        uint32_t b = __float_as_uint(v);
        return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
    }
    __device__ __forceinline__ static float dec(uint32_t k) {
        uint32_t b = (k >> 31) ? (k ^ 0x80000000u) : ~k;
        return __uint_as_float(b);
    }
};
template <> struct KeyOf<double> {
    typedef uint64_t type;
    static constexpr int bits = 64;
    __device__ __forceinline__ static uint64_t enc(double v) {
        uint64_t b = (uint64_t)__double_as_longlong(v);
        return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
    }
    __device__ __forceinline__ static double dec(uint64_t k) {
        uint64_t b = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
        return __longlong_as_double((long long)b);
    }
};

// ---------------------------------------------------------------------------------------------
// Walk state, resident in the workspace.  Written only by single-thread sections of kernels.
// ---------------------------------------------------------------------------------------------
struct WalkState {
    // parameters (crwr_walk_init)
    double rp, perc, tol;
    int32_t max_steps;
    // progress
    int32_t step;            // index of the step being / about to be executed
    int32_t done;            // stop rule fired
    int32_t stop_reason;
    int32_t plateau;         // utility.py:298-299 cumulative counter
    double delta_prev;       // utility.py:301
    int32_t parity;          // buffer holding the iterate the next propagate reads
    int32_t cut_on;          // 1: entries <= cut are dropped when the iterate is read (lazy cut)
    double cut;              // value of the last applied cut (exact in the walk dtype)
    // per-step counters, reset by the finish kernel
    unsigned long long cursor;      // entries allocated in the output buffer (= local nnz_new)
    unsigned long long nnz_in;      // kept entries read this step (local)
    unsigned long long nnz_exact;   // stored (non-zero) entries of the step's product (cursor may include zero padding)
    unsigned int row_counter;       // dynamic row scheduler of the main propagate kernel
    unsigned int fb_counter;        // ... of the large-row kernel
    unsigned int fb_count;          // rows queued for the large-row kernel
    int32_t max_row_len;
    int32_t max_kept;               // longest kept list read this step
    int32_t last_max_kept;
    int32_t last_max_row_len;
    int32_t pad1;
    int32_t capacity_overflow;      // sticky
    int32_t select_overflow;        // sticky: candidate buffer overflow
    // select
    unsigned long long n_total;     // global number of values (sum of level-0 histogram)
    unsigned long long rank_k;      // floor of the virtual index
    unsigned long long below;       // values strictly below the current prefix bin
    unsigned long long prefix;      // accumulated digits (level 0 in the high bits)
    unsigned long long bin_count;   // population of the selected bin
    double gamma;                   // fractional part of the virtual index (walk dtype precision)
    int32_t clamp_last;             // virtual index >= n-1: both order statistics are the maximum
    int32_t last_fallback_rows;
    double last_delta;
    unsigned long long sq_limb[3];  // sum of per-row ||dQ||^2 in 2^-64 fixed point, 32-bit limbs
    // fused select (unsharded walks): histograms are filled by the propagate epilogue, scans and the
    // finish step run in the last CTA of the following kernels
    unsigned long long pred_prefix0;   // level-0 bin of the previous cut: level-1 histogram is pre-filled for it
    int32_t pred_valid;                // pred_prefix0 is meaningful
    int32_t hist1_ready;               // level-1 histogram already holds the counts of the selected level-0 bin
    unsigned int tick_dense, tick_hist1, tick_gather;   // last-CTA tickets
    unsigned int tick_prop;
    int32_t peer_timeout;              // a cross-GPU barrier timed out (a peer rank is gone): the walk stops
    unsigned int slow2_count;          // rows the warp-per-row symbolic kernel handed to the CTA kernel this step
    int32_t debug_error;               // CRWR_DEBUG_CHECKS builds: highest failed bounds-check code (0 = none)
    int32_t order_hazard;              // sticky: unordered rows would have been consumed without a cut (never for positive data)
    int32_t scan_done;                 // the level scans of this step already ran (in the propagate kernel's last CTA)
    int32_t unordered_rows;            // the row-group kernel wrote this step's rows in slot-claim order (valid only behind a cut)
    // cached row structures (structure.cuh): rows whose kept set lies inside the cached superset take the fast kernel
    unsigned long long pool_cursor;    // bytes of the structure pool in use
    unsigned long long slot_cursor;    // entries of the fixed-slot region (both iterate buffers) in use
    unsigned long long last_nnz_new;   // stored entries of the last finished step's product (all ranks)
    unsigned long long hot_sum;        // bytes of the hot parts of all current structures (mod 2^64; the host sizes the fast kernel's stages)
    unsigned int pool_gen;             // generation of this walk's structures (0 in a header = no structure)
    int32_t struct_full;               // sticky: slot region or pool exhausted, later rows went without a structure
    unsigned int slow_count;           // rows the fast kernel handed to the row-group kernel this step
    unsigned int fast_rows;            // rows the fast kernels computed this step
    unsigned int big_rows;             // long rows among them / built this step (the host then queues the big-row kernel)
    int32_t win_narrow;                // the window set for the next step is narrow enough for a sharded walk's push regions
    unsigned int built_rows;           // structures built this step
    int32_t fb_unhandled;              // sticky: rows were deferred in a step that launched no large-row kernel (host retries)
    // window select (tail.cuh): the epilogues collect the values in [win_lo, win_hi) as percentile candidates
    int32_t win_on;                    // the current step's epilogues classify against the window
    unsigned long long win_lo, win_hi; // window bounds as order-preserving keys
    unsigned long long win_below;      // values strictly below win_lo (this rank)
    unsigned int win_miss_steps;       // diagnostics: steps whose window missed (full select in the tail kernel)
    int32_t win_missed;                // sharded walk: the window missed, the step was rewound and the walk waits for the host
    double win_prev_rel;               // relative move of the cut one step earlier (the window follows the larger of the last two)
};

// Exchange block 0 of a sharded walk: level-0 histogram followed by these extras (all uint64, SUM).
constexpr int kXSq0 = 0, kXSq1 = 1, kXSq2 = 2, kXNnzIn = 3, kXNnzNew = 4, kXCapOvf = 5, kXSelOvf = 6, kXFallback = 7;
constexpr int kXExtras = 8;
// Candidate block: [count, successor key, keys...] (uint64).
constexpr int kCandHeader = 2;

// Bounds checks of our own (compute-sanitizer is closed on this pool): a -DCRWR_DEBUG_CHECKS build
// records the highest failed check code in the walk state instead of trapping; crwr_walk_poll reports it.
#ifdef CRWR_DEBUG_CHECKS
#define CRWR_CHECK(st, cond, code) do { if (!(cond)) atomicMax(&(st)->debug_error, (int)(code)); } while (0)
// loop guard: after `limit` iterations the code is recorded and `action` leaves the loop
#define CRWR_GUARD_DECL(g) int g = 0
#define CRWR_GUARD(st, g, limit, code, action) do { if (++(g) > (limit)) { atomicMax(&(st)->debug_error, (int)(code)); action; } } while (0)
#else
#define CRWR_GUARD_DECL(g) do { } while (0)
#define CRWR_GUARD(st, g, limit, code, action) do { } while (0)
#define CRWR_CHECK(st, cond, code) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Exact, partition-independent sums of squares (||Q - Q~||_F^2, utility.py:283): every term is converted
// to 2^-64 fixed point on its own (integer part and fraction kept apart), so the integer sums do not
// depend on which lanes, warps, kernels or GPUs added them.
// ---------------------------------------------------------------------------------------------
struct Sq128 { unsigned long long lo, hi; };
__device__ __forceinline__ void sq_add(Sq128& a, double x) {      // x >= 0
    if (!(x > 0.0)) return;
    if (x < 1.0) {            // the usual case; same bits as the general path (integer part 0)
        const unsigned long long f = __double2ull_rd(x * 18446744073709551616.0);
        a.lo += f;
        a.hi += (a.lo < f) ? 1ull : 0ull;
        return;
    }
    if (x >= 9.0e18) x = 9.0e18;
    const double h = floor(x);
    const unsigned long long f = __double2ull_rd((x - h) * 18446744073709551616.0);
    a.hi += __double2ull_rd(h);
    a.lo += f;
    a.hi += (a.lo < f) ? 1ull : 0ull;
}
__device__ __forceinline__ void sq_merge(Sq128& a, unsigned long long lo, unsigned long long hi) {
    a.lo += lo;
    a.hi += hi + ((a.lo < lo) ? 1ull : 0ull);
}
// Sum over the lanes of an aligned group of `width` lanes (all lanes of the warp take part).
__device__ __forceinline__ Sq128 sq_group_sum(Sq128 s, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) {
        const unsigned long long olo = __shfl_xor_sync(kFull, s.lo, o);
        const unsigned long long ohi = __shfl_xor_sync(kFull, s.hi, o);
        sq_merge(s, olo, ohi);
    }
    return s;
}
// pos - neg, where neg's terms are a subset of pos's (so the difference is a sum of squares again)
__device__ __forceinline__ Sq128 sq_diff(const Sq128& pos, const Sq128& neg) {
    Sq128 r;
    r.lo = pos.lo - neg.lo;
    r.hi = pos.hi - neg.hi - ((pos.lo < neg.lo) ? 1ull : 0ull);
    return r;
}


// Cached symbolic structure of one restart row (structure.cuh).
struct RowStruct {
    long long off;          // byte offset of the row's block in the structure pool
    long long slot;         // the row's fixed slot: first entry in both iterate buffers (beyond the per-step region)
    int32_t no;             // output columns (= entries of the slot)
    int32_t nprod;          // products
    int32_t nk;             // |K+|: operand columns the structure covers
    int32_t dpos;           // position of the restart column among the outputs
    uint32_t gen;           // walk generation the block belongs to; 0 = no structure
    int32_t step;           // last walk step that wrote the row in structure order into its slot
    int32_t hot_bytes;      // bytes of the block's hot part (what the fast kernel stages in shared memory)
    int32_t groups;         // groups the outputs are dealt to (1, or kBigGroups for a long row)
    int32_t blk_cap;        // bytes the row's pool block can hold (a rebuilt structure that fits stays in place)
    int32_t slot_cap;       // entries the row's fixed slot can hold
};

template <typename T>
struct IterBuf {
    int32_t* cols;
    T* vals;
    long long* row_ptr;
    int32_t* row_len;
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): the kernels of a walk step are launched with programmatic stream
// serialization, so the CTAs of kernel N+1 are brought onto the SMs while kernel N drains and only wait here
// until N has completed and its writes are visible - the launch gap between the dependent kernels of a step
// disappears.  Both are no-ops for a kernel launched the ordinary way.  Rule: every kernel of the chain waits
// before it reads anything another kernel wrote and before it exits (completion is transitive only then).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <typename K> __device__ __forceinline__ K warp_min(K v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        K w = __shfl_xor_sync(kFull, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// Ascending sort of (key[i], val[i]) for i < n by one warp; all-ascending bitonic network with
// "flip" first stages, valid for any n (out-of-range partners act as +inf).  Keys are unique.
template <typename K, typename V>
__device__ __forceinline__ void warp_sort_pairs(K* key, V* val, int n) {
    const int lane = lane_id();
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = lane; i < n; i += 32) {
            int l = i ^ (k - 1);
            if (l > i && l < n) {
                K a = key[i], b = key[l];
                if (b < a) { key[i] = b; key[l] = a; V t = val[i]; val[i] = val[l]; val[l] = t; }
            }
        }
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                int l = i ^ j;
                if (l > i && l < n) {
                    K a = key[i], b = key[l];
                    if (b < a) { key[i] = b; key[l] = a; V t = val[i]; val[i] = val[l]; val[l] = t; }
                }
            }
            __syncwarp();
        }
    }
}

// Same, for n <= 32 pairs living in global memory: one pair per lane, sorted in registers with shuffles
// (the memory-resident network above costs a round trip per stage when the arrays are in global memory).
template <typename V>
__device__ __forceinline__ void warp_sort_pairs32(int32_t* key, V* val, int n) {
    const int lane = lane_id();
    int32_t k = lane < n ? key[lane] : 0x7fffffff;
    V v = lane < n ? val[lane] : (V)0;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
            const int32_t ok = __shfl_xor_sync(kFull, k, j);
            const V ov = __shfl_xor_sync(kFull, v, j);
            const bool asc = (lane & size) == 0;
            const bool lower = (lane & j) == 0;
            const bool take_other = (lower == asc) ? (ok < k) : (ok > k);     // keys are unique (padding excepted)
            if (take_other) { k = ok; v = ov; }
        }
    }
    if (lane < n) { key[lane] = k; val[lane] = v; }
}

template <typename K, typename V>
__device__ __forceinline__ void warp_sort_pairs_any(K* key, V* val, int n) {
    if (n <= 32) warp_sort_pairs32(key, val, n);
    else warp_sort_pairs(key, val, n);
}

// numpy's pairwise summation as np.add.reduceat applies it to one contiguous segment:
// x[0] + pairwise(x[1:])  (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE 128).
template <typename T>
__device__ T np_pairwise(const T* x, long long n) {
    if (n < 8) {
        T r = (T)0;
        for (long long i = 0; i < n; ++i) r = add_rn(r, x[i]);
        return r;
    } else if (n <= 128) {
        T r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = x[j];
        long long i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = add_rn(r[j], x[i + j]);
        }
        T res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])), add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
        for (; i < n; ++i) res = add_rn(res, x[i]);
        return res;
    } else {
        long long n2 = n / 2;
        n2 -= n2 % 8;
        return add_rn(np_pairwise<T>(x, n2), np_pairwise<T>(x + n2, n - n2));
    }
}
template <typename T>
__device__ T np_segment_sum(const T* x, long long n) {   // n >= 1
    if (n == 1) return x[0];
    return add_rn(x[0], np_pairwise<T>(x + 1, n - 1));
}

}  // namespace crwr
