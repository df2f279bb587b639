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
#include "select.cuh"

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
    unsigned long long* hist;   // [2][4096] select histograms; null: do not fuse (sharded walk, step 0)
    int32_t fused_select;       // 1: the last CTA of the large-row kernel scans the histograms
    int32_t chunk;              // entries a warp reserves from the product buffer at a time
    int32_t row_chunk;          // rows a warp takes from the row counter at a time
    int32_t prefill_l1;         // 1: also count level 1 for the level-0 bin predicted from the previous cut
    const int32_t* row_order;   // row-group kernel: rows in the order they are handed out (longest first / slow list), or null
    // cached row structures and window select (structure.cuh, tail.cuh)
    const unsigned int* row_list_count;   // device count of the rows listed in row_order (slow list), or null: n_rows
    RowStruct* rs;              // [n_rows] structure headers
    unsigned char* pool;        // structure pool
    unsigned long long pool_cap;
    int32_t* slow_list;         // fast kernel: rows it could not take
    float alpha;                // K+ = entries above alpha x cut
    long long fixed_base;       // first entry of the fixed-slot region of both buffers (behind the out_cap per-step entries)
    long long slot_cap;         // its entries
    int32_t slot_chunk;         // symbolic kernel: entries / bytes a warp reserves at a time while many rows are built
    int32_t pool_chunk;
    unsigned char* sym_tmp;     // symbolic kernel: per-CTA staging regions (rows too long for shared-memory staging)
    int32_t small_no, small_hot;   // symbolic kernel: what a stage of the fast kernel holds (outputs, bytes); longer rows get groups
    int32_t win_mode;           // classify values against the select window when the walk state has one
    unsigned long long* cand;   // candidate block [count, successor, keys...] of the window select
    unsigned long long cand_cap;
};

// Fused select, unsharded walk: once every value of the step is in the level-0 histogram, one CTA locates
// the level-0 bin of the percentile and, when the level-1 histogram was pre-filled for that very bin
// (prediction from the previous cut), the 24-bit prefix as well; otherwise level 1 is cleared for the
// histogram pass that follows.  Called by all BLOCK threads of the last CTA.
template <typename T, int BLOCK>
__device__ __noinline__ void scan_tail(WalkState* st, unsigned long long* hist) {
    scan_level<T, BLOCK>(st, hist, 0);
    const bool hit = st->pred_valid && st->prefix == st->pred_prefix0;
    __syncthreads();
    if (hit) {
        scan_level<T, BLOCK>(st, hist, 1);
        if (threadIdx.x == 0) { st->hist1_ready = 1; st->scan_done = 1; }
    } else {
        for (int i = threadIdx.x; i < kSelectBins; i += BLOCK) hist[kSelectBins + i] = 0ull;
        if (threadIdx.x == 0) { st->hist1_ready = 0; st->scan_done = 1; }
    }
}

// Per-warp running totals of a launch.  Everything a row contributes to the walk state is summed here
// (identical values in all lanes) and published with a handful of atomics when the warp retires:
// same-address atomics from every row serialise in L2 and were the bottleneck of small-row workloads.
struct WarpTotals {
    unsigned long long sq0, sq1, sq2;   // ||dQ||^2 as three 32-bit limbs of a 2^-64 fixed-point number
    unsigned long long nnz_exact, kept;
    int max_kept, max_row_len;
    // output space: a private chunk of the product buffer, refilled from the global cursor
    long long ch_base;
    int ch_left;
    bool ovf;
};

__device__ __forceinline__ void totals_init(WarpTotals& w) {
    w.sq0 = w.sq1 = w.sq2 = w.nnz_exact = w.kept = 0ull;
    w.max_kept = w.max_row_len = 0;
    w.ch_base = 0; w.ch_left = 0; w.ovf = false;
}

// Integer limbs commute, so delta is bit-identical whatever the order in which rows finish and however
// the rows are sharded over GPUs.
__device__ __forceinline__ void totals_add_rowsq(WarpTotals& w, const Sq128& r) {
    w.sq2 += r.hi;
    w.sq1 += r.lo >> 32;
    w.sq0 += r.lo & 0xffffffffull;
}

template <typename T>
__device__ __forceinline__ void pad_zero(const IterBuf<T>& out, long long base, int n) {
    for (int i = lane_id(); i < n; i += 32) { out.cols[base + i] = -1; out.vals[base + i] = (T)0; }
}

// Reserves n entries of the product buffer for a row (uniform across the warp).  Space comes from the
// warp's private chunk; an exhausted chunk is zero-padded (every pass over the value array ignores
// zeros: a stored value is never zero) and replaced through ONE atomic on the global cursor.
template <typename T>
__device__ __forceinline__ long long reserve_row(const PropArgs<T>& a, WarpTotals& w, int n) {
    if (n > w.ch_left) {
        if (w.ch_left > 0 && !w.ovf) pad_zero(a.out, w.ch_base, w.ch_left);
        const int sz = n > a.chunk ? n : a.chunk;
        long long base = 0;
        if (lane_id() == 0) base = (long long)atomicAdd(&a.st->cursor, (unsigned long long)sz);
        base = __shfl_sync(kFull, base, 0);
        w.ch_base = base;
        w.ch_left = sz;
        if (base + sz > a.out_cap) {
            w.ovf = true;
            if (lane_id() == 0) a.st->capacity_overflow = 1;
        }
    }
    const long long r = w.ch_base;
    w.ch_base += n;
    w.ch_left -= n;
    return r;
}

// Writes one finished row into its reserved space.  `emit(i, &col, &val)` yields the i-th candidate
// (i < n_cand) in output order; zero values are dropped (csr_matmat / csr_plus_csr never store them) and
// the unused tail of the reservation is zero-padded.
template <typename T, typename Emit>
__device__ __forceinline__ void write_row(const PropArgs<T>& a, WarpTotals& w, int row, long long base, int n_cand,
                                          int n_nonzero, Emit emit) {
    const int lane = lane_id();
    const bool fits = !w.ovf;
    if (lane == 0) {
        a.out.row_ptr[row] = fits ? base : 0;
        a.out.row_len[row] = fits ? n_nonzero : 0;
    }
    w.max_row_len = max(w.max_row_len, n_nonzero);
    w.nnz_exact += (unsigned long long)n_nonzero;
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
            CRWR_CHECK(a.st, o >= 0 && o < a.out_cap, 6);
            a.out.cols[o] = c;
            a.out.vals[o] = v;
        }
        written += __popc(m);
    }
    if (written < n_cand) pad_zero(a.out, base + written, n_cand - written);
}

// Retires a warp: pads its unused chunk and publishes its totals.
template <typename T>
__device__ __forceinline__ void totals_flush(const PropArgs<T>& a, WarpTotals& w) {
    if (w.ch_left > 0 && !w.ovf) pad_zero(a.out, w.ch_base, w.ch_left);
    if (lane_id() == 0) {
        WalkState* st = a.st;
        if (w.sq0) atomicAdd(&st->sq_limb[0], w.sq0);
        if (w.sq1) atomicAdd(&st->sq_limb[1], w.sq1);
        if (w.sq2) atomicAdd(&st->sq_limb[2], w.sq2);
        if (w.nnz_exact) atomicAdd(&st->nnz_exact, w.nnz_exact);
        if (w.kept) atomicAdd(&st->nnz_in, w.kept);
        if (w.max_kept) atomicMax(&st->max_kept, w.max_kept);
        if (w.max_row_len) atomicMax(&st->max_row_len, w.max_row_len);
    }
}

// Rows are handed out `chunk` at a time by an atomic counter.
__device__ __forceinline__ bool next_row(unsigned int* counter, int n_rows, int chunk, int& cur, int& lim) {
    if (cur == lim) {
        int r = 0;
        if (lane_id() == 0) r = (int)atomicAdd(counter, (unsigned)chunk);
        r = __shfl_sync(kFull, r, 0);
        cur = r;
        lim = min(r + chunk, n_rows);
    }
    return cur < n_rows;
}

// Fused level-0 histogram of the percentile select (and level 1 for the predicted level-0 bin).
// Level 0: values are counted into a per-CTA shared-memory window of kHistWin bins (one ATOMS per
// distinct bin among a chunk's 32 values) that is flushed once at the end of the kernel - global
// same-address atomics cost ~8 ns each on the hot bins.  Level 1 only concerns ~5% of the values and
// spreads them over 4096 addresses, so it goes to global memory directly.
template <typename T>
__device__ __forceinline__ void hist_values(unsigned long long* hist, unsigned int* win, bool valid, T v,
                                            bool pred_valid, typename KeyOf<T>::type pred0) {
    typedef KeyOf<T> KO;
    const typename KO::type key = KO::enc(v);
    const unsigned bin0 = valid ? (unsigned)(key >> (KO::bits - kSelectBits)) : (0x80000000u | lane_id());
    const unsigned peers = __match_any_sync(kFull, bin0);
    if (valid && (peers & lanemask_lt()) == 0u) {
        const unsigned w = bin0 - (unsigned)HistWindow<T>::base;
        if (w < (unsigned)kHistWin) atomicAdd(&win[w], (unsigned)__popc(peers));
        else atomicAdd(&hist[bin0 & (kSelectBins - 1)], (unsigned long long)__popc(peers));
    }
    if (valid && pred_valid && (key >> (KO::bits - kSelectBits)) == pred0)
        atomicAdd(&hist[kSelectBins + ((unsigned)(key >> (KO::bits - 2 * kSelectBits)) & (kSelectBins - 1))], 1ull);
}

template <typename T>
__device__ __forceinline__ void hist_window_flush(unsigned long long* hist, const unsigned int* win) {
    for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) {
        const unsigned c = win[i];
        if (c) atomicAdd(&hist[HistWindow<T>::base + i], (unsigned long long)c);
    }
}

constexpr int kWinStage = 256;             // window candidates staged per CTA (shares the histogram window's shared memory)

// ---------------------------------------------------------------------------------------------
// Window select, producer side (consumer: tail.cuh).  Once the cut has settled the next cut is known to lie in
// a narrow window [lo, hi) around the previous one; the epilogues then only count the values below the window,
// collect the few inside it as candidates and track the smallest value at or above it - the whole-array
// histogram and gather passes of the radix select disappear.  Candidates are staged per CTA in shared memory
// and appended to the candidate block [count, successor, keys...] with one atomic per CTA.
// ---------------------------------------------------------------------------------------------
struct WinAcc {
    unsigned long long below;       // values below the window seen by this thread
    unsigned long long succ;        // smallest key >= hi seen by this thread
};
struct WinStage {
    unsigned long long* keys;       // [kWinStage] shared memory
    unsigned int* count;            // shared
};
template <typename T>
__device__ __forceinline__ void win_classify(bool valid, T v, unsigned long long lo, unsigned long long hi, WinAcc& w,
                                             const WinStage& s, unsigned long long* cand, unsigned long long cand_cap) {
    if (!valid) return;
    const unsigned long long key = (unsigned long long)KeyOf<T>::enc(v);
    if (key < lo) { w.below += 1; return; }
    if (key >= hi) { w.succ = key < w.succ ? key : w.succ; return; }
    const unsigned int pos = atomicAdd(s.count, 1u);
    if (pos < (unsigned)kWinStage) s.keys[pos] = key;
    else {
        const unsigned long long idx = atomicAdd(&cand[0], 1ull);
        if (idx < cand_cap) cand[kCandHeader + idx] = key;
    }
}
// CTA-wide, after a barrier that follows the last win_classify.
__device__ __forceinline__ void win_flush(const WinAcc& w, const WinStage& s, unsigned long long* cand, unsigned long long cand_cap,
                                          WalkState* st) {
    __shared__ unsigned long long s_base, s_below, s_succ;
    if (threadIdx.x == 0) { s_below = 0ull; s_succ = ~0ull; }
    __syncthreads();
    unsigned long long b = w.below, m = w.succ;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b += __shfl_xor_sync(kFull, b, o);
        const unsigned long long om = __shfl_xor_sync(kFull, m, o);
        m = om < m ? om : m;
    }
    if (lane_id() == 0) {
        if (b) atomicAdd(&s_below, b);
        if (m != ~0ull) atomicMin(&s_succ, m);
    }
    const unsigned int n = *s.count < (unsigned)kWinStage ? *s.count : (unsigned)kWinStage;
    if (threadIdx.x == 0 && n) s_base = atomicAdd(&cand[0], (unsigned long long)n);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_below) atomicAdd(&st->win_below, s_below);
        if (s_succ != ~0ull) atomicMin(&cand[1], s_succ);
    }
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long o = s_base + i;
        if (o < cand_cap) cand[kCandHeader + o] = s.keys[i];
    }
}

__device__ __forceinline__ uint32_t hash_col(int32_t j, int log_h) {
    return ((uint32_t)j * 0x9E3779B1u) >> (32 - log_h);
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
    const size_t ht = (size_t)H + 64;
    return align16(ht * sizeof(T)) + align16((size_t)kmax * sizeof(T)) + 2 * align16((size_t)2 * scap * sizeof(T)) +
           align16(ht * 4) + 3 * align16((size_t)kmax * 4) + align16((size_t)2 * scap * 4) +
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

// Ascending sort of up to 256 (column, value) pairs held in shared memory, by one warp, in
// registers: 8 keys per lane, bitonic network with shuffles.  Keys are (column << 32 | position);
// values are gathered afterwards through the positions.
template <typename T, int R>
__device__ __forceinline__ void warp_sort_kept_regs(int32_t* kcol, T* kval, T* scratch, int n) {
    const int lane = lane_id();
    unsigned long long key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;          // element i lives in register r of lane (i & 31)
        key[r] = i < n ? (((unsigned long long)(uint32_t)kcol[i] << 32) | (unsigned)i) : ~0ull;
    }
    // index of element = r*32 + lane: bit positions 0..4 = lane bits, 5..7 = register bits
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rj = j >> 5;        // partner differs in a register bit
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const int i = r * 32 + lane;
                        const bool asc = (i & k) == 0;
                        unsigned long long x = key[r], y = key[r | rj];
                        const bool sw = (x > y) == asc;
                        key[r] = sw ? y : x;
                        key[r | rj] = sw ? x : y;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = r * 32 + lane;
                    const bool asc = (i & k) == 0;
                    const unsigned long long other = __shfl_xor_sync(kFull, key[r], j);
                    const bool lower = (lane & j) == 0;
                    // the lower index keeps the min when ascending
                    const bool take_min = lower == asc;
                    key[r] = take_min ? (key[r] < other ? key[r] : other) : (key[r] > other ? key[r] : other);
                }
            }
        }
    }
    // gather values through the sorted positions (scratch holds a copy of the unsorted values)
    for (int i = lane; i < n; i += 32) scratch[i] = kval[i];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        if (i < n) {
            kcol[i] = (int32_t)(key[r] >> 32);
            kval[i] = scratch[(unsigned)(key[r] & 0xffffffffu)];
        }
    }
    __syncwarp();
}

// One walk step's propagation for the rows of this shard (all CTAs of the grid cooperate through the
// row counter).  Called by every thread of every CTA; smem_raw is the CTA's dynamic shared memory.
// COHERENT: read the operand iterate with ordinary loads (required inside the persistent kernel, which
// rewrites that buffer two steps later); otherwise through the read-only path.
template <typename T, bool COHERENT>
__device__ __forceinline__ void propagate_rows(const PropArgs<T>& a, const PropShape& sh, unsigned char* smem_raw) {
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int H = sh.H, KMAX = sh.kmax, SCAP = sh.scap, LIMIT = sh.limit;
    // table of H slots plus 64 slots of probe overrun; the last one is a never-claimed sentinel that
    // stops a probe and sends it back to slot 0 (so the probe loop itself needs no wrap test)
    const int HT = H + 64;
    unsigned char* p = smem_raw + (size_t)warp * sh.per_warp;
    T* tval = reinterpret_cast<T*>(p);            p += align16((size_t)HT * sizeof(T));
    T* kval = reinterpret_cast<T*>(p);            p += align16((size_t)KMAX * sizeof(T));
    T* sval = reinterpret_cast<T*>(p);            p += align16((size_t)2 * SCAP * sizeof(T));
    T* sq = reinterpret_cast<T*>(p);              p += align16((size_t)2 * SCAP * sizeof(T));
    int32_t* tkey = reinterpret_cast<int32_t*>(p); p += align16((size_t)HT * 4);
    int32_t* kcol = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* kps = reinterpret_cast<int32_t*>(p);  p += align16((size_t)KMAX * 4);
    int32_t* klen = reinterpret_cast<int32_t*>(p); p += align16((size_t)KMAX * 4);
    int32_t* scol = reinterpret_cast<int32_t*>(p); p += align16((size_t)2 * SCAP * 4);
    uint16_t* ord = reinterpret_cast<uint16_t*>(p);

    unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw + (size_t)sh.warps * sh.per_warp);
    for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) win[i] = 0u;
    for (int i = lane; i < HT; i += 32) tkey[i] = -1;
    __syncthreads();

    const bool cut_on = __ldcg(&a.st->cut_on) != 0;
    const T cut = (T)__ldcg(&a.st->cut);
    const bool pred_valid = a.prefill_l1 && __ldcg(&a.st->pred_valid) != 0;
    const typename KeyOf<T>::type pred0 = (typename KeyOf<T>::type)__ldcg(&a.st->pred_prefix0);
    const int32_t* __restrict__ pcols = a.p_cols;
    const T* __restrict__ pvals = a.p_vals;

    // Copies the P rows of up to 32 operand entries (lane t <-> entry first+t, each lane streams its
    // own row with cp.async) into staging buffer `buf`, together with the entry's value q.  Returns
    // the number of entries taken (the longest prefix whose products fit) and the product count.
    auto stage = [&](int first, int nk, int buf, int& n_prod) -> int {
        n_prod = 0;
        if (first >= nk) return 0;
        const int idx = first + lane;
        const int len = idx < nk ? klen[idx] : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned fit = __ballot_sync(kFull, idx < nk && incl <= SCAP);
        int nb = fit == kFull ? 32 : __ffs(~fit) - 1;
        int32_t* dc = scol + buf * SCAP;
        T* dv = sval + buf * SCAP;
        T* dq = sq + buf * SCAP;
        if (nb == 0) {
            // a single P row longer than the staging buffer: its first SCAP entries now, the rest by
            // re-entering with the extent advanced (kps/klen of the entry are updated in place)
            const int ps0 = kps[first];
            const T q0 = kval[first];
            for (int e = lane; e < SCAP; e += 32) {
                cp_async4(dc + e, pcols + ps0 + e);
                cp_async_val(dv + e, pvals + ps0 + e);
                dq[e] = q0;
            }
            __syncwarp();
            if (lane == 0) { kps[first] = ps0 + SCAP; klen[first] -= SCAP; }
            __syncwarp();
            n_prod = SCAP;
            return 0;                      // entry not finished: `first` does not advance
        }
        if (lane < nb) {
            const int off = incl - len;
            const int ps = kps[idx];
            const T q = kval[idx];
            const int32_t* gc = pcols + ps;
            const T* gv = pvals + ps;
            CRWR_CHECK(a.st, off + len <= SCAP, 3);
            for (int e = 0; e < len; ++e) {
                cp_async4(dc + off + e, gc + e);
                cp_async_val(dv + off + e, gv + e);
                dq[off + e] = q;
            }
        }
        n_prod = __shfl_sync(kFull, incl, nb - 1);
        return nb;
    };

    WarpTotals wt;
    totals_init(wt);
    int row_cur = 0, row_lim = 0;
    for (;;) {
        if (!next_row(&a.st->row_counter, a.n_rows, a.row_chunk, row_cur, row_lim)) break;
        const int row = row_cur++;

        // ---- read the row, dropping entries at or below the previous cut (strict >, utility.py:290)
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        const int32_t* rc = a.in.cols + ptr;      // plain (coherent) loads: the persistent walk kernel rewrites
        const T* rv = a.in.vals + ptr;            // this buffer two steps later
        int nk = 0;
        for (int b = 0; b < len; b += 128) {
            int32_t c[4];
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * 32 + lane;
                c[u] = 0; v[u] = (T)0;
                if (e < len) {
                    if (COHERENT) { c[u] = rc[e]; v[u] = rv[e]; }
                    else { c[u] = __ldg(rc + e); v[u] = __ldg(rv + e); }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * 32 + lane;
                const bool keep = (e < len) && (!cut_on || v[u] > cut);
                const unsigned m = __ballot_sync(kFull, keep);
                const int pos = nk + __popc(m & lanemask_lt());
                if (keep && pos < KMAX) { CRWR_CHECK(a.st, c[u] >= 0, 4); kcol[pos] = c[u]; kval[pos] = v[u]; }
                nk += __popc(m);
            }
        }
        __syncwarp();
        wt.max_kept = max(wt.max_kept, nk);
        bool defer = nk > KMAX;
        int count = 0;
        if (!defer) {
            if (cut_on && nk > 1) {
                if (nk <= 32) warp_sort_kept_regs<T, 1>(kcol, kval, sval, nk);
                else if (nk <= 64) warp_sort_kept_regs<T, 2>(kcol, kval, sval, nk);
                else if (nk <= 128) warp_sort_kept_regs<T, 4>(kcol, kval, sval, nk);
                else if (nk <= 256) warp_sort_kept_regs<T, 8>(kcol, kval, sval, nk);
                else warp_sort_pairs(kcol, kval, nk);
            }
            for (int idx = lane; idx < nk; idx += 32) {
                const int k = kcol[idx];
                CRWR_CHECK(a.st, k >= 0, 5);
                const int ps = __ldg(a.p_indptr + k);
                kps[idx] = ps;
                klen[idx] = __ldg(a.p_indptr + k + 1) - ps;
            }
            __syncwarp();

            // ---- the row's products in scipy's order (k ascending, then P-row order), 32 per round;
            // batch b+1 is copied into shared memory (cp.async) while batch b is accumulated.
            int first = 0, buf = 0, n_cur = 0, n_nxt = 0;
            int nb_cur = stage(first, nk, buf, n_cur);
            cp_async_commit();
            while (n_cur > 0 || nb_cur > 0) {
                const int next_first = first + nb_cur;
                const int nb_nxt = stage(next_first, nk, buf ^ 1, n_nxt);
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const int32_t* bc = scol + buf * SCAP;
                const T* bv = sval + buf * SCAP;
                const T* bq = sq + buf * SCAP;
                // round r+1's products and their duplicate groups (MATCH) are fetched while round r probes
                int32_t j_n = -1 - lane;
                T prod_n = (T)0;
                if (lane < n_cur) { j_n = bc[lane]; prod_n = mul_rn(bq[lane], bv[lane]); }
                unsigned peers_n = __match_any_sync(kFull, j_n);
                for (int r0 = 0; r0 < n_cur; r0 += 32) {
                    if (count + 32 > LIMIT) { defer = true; break; }
                    const int32_t j = j_n;
                    const T prod = prod_n;
                    const unsigned peers = peers_n;
                    const bool active = r0 + lane < n_cur;
                    {
                        const int fn = r0 + 32 + lane;
                        j_n = -1 - lane;
                        prod_n = (T)0;
                        if (fn < n_cur) { j_n = bc[fn]; prod_n = mul_rn(bq[fn], bv[fn]); }
                        peers_n = __match_any_sync(kFull, j_n);
                    }
                    // lanes holding the same column: the lowest one (earliest product) owns the slot and
                    // adds the others' products in lane order = scipy's order
                    const bool leader = active && ((peers & lanemask_lt()) == 0u);
                    const bool dup = __any_sync(kFull, (peers & (peers - 1u)) != 0u);
                    bool is_new = false;
                    uint32_t slot = 0;
                    T acc = prod;
                    if (leader) {
                        slot = hash_slot(j, H);
                        for (;;) {
                            int32_t kk = tkey[slot];
                            while (kk != j && kk != -1) { ++slot; CRWR_CHECK(a.st, slot < (uint32_t)HT, 1); kk = tkey[slot]; }   // read-only probe
                            if (kk == j) { acc = add_rn(tval[slot], prod); break; }
                            if (slot == (uint32_t)(HT - 1)) { slot = 0; continue; }   // sentinel slot: wrap
                            if (atomicCAS(&tkey[slot], -1, j) == -1) { is_new = true; break; }
                            // lost the empty slot to another column of this round: keep probing
                        }
                    }
                    if (dup) {
                        unsigned rem = peers & (peers - 1u);       // peers after the leader
                        const int maxmult = __reduce_max_sync(kFull, __popc(peers));
                        for (int s_ = 1; s_ < maxmult; ++s_) {
                            const int src = rem ? (__ffs(rem) - 1) : lane;
                            const T other = __shfl_sync(kFull, prod, src);
                            if (leader && rem) acc = add_rn(acc, other);
                            rem &= rem - 1u;
                        }
                    }
                    if (leader) tval[slot] = acc;
                    const unsigned nm = __ballot_sync(kFull, is_new);
                    if (nm) {
                        CRWR_CHECK(a.st, count + 32 <= LIMIT + 32, 2);
                        if (is_new) ord[count + __popc(nm & lanemask_lt())] = (uint16_t)slot;
                        count += __popc(nm);
                    }
                    __syncwarp();
                }
                if (defer) break;
                first = next_first;
                buf ^= 1;
                nb_cur = nb_nxt;
                n_cur = n_nxt;
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
        wt.kept += (unsigned long long)nk;

        // ---- epilogue: scale, restart add, norm, write
        const int32_t dcol = a.row_begin + row;
        int dslot = -1;
        {
            uint32_t slot = hash_slot(dcol, H);
            for (;;) {
                const int32_t kk = tkey[slot];
                if (kk == dcol) { dslot = (int)slot; break; }
                if (kk == -1) {
                    if (slot == (uint32_t)(HT - 1)) { slot = 0; continue; }
                    break;
                }
                ++slot;
            }
        }
        const bool diag_new = dslot < 0;
        const int n_cand = count + (diag_new ? 1 : 0);
        const long long base = reserve_row(a, wt, n_cand);
        Sq128 pos = {0ull, 0ull}, neg = {0ull, 0ull};
        int nz = 0;
        for (int b = 0; b < count; b += 32) {
            const int i = b + lane;
            bool one = false;
            T v = (T)0;
            if (i < count) {
                const int slot = ord[i];
                v = mul_rn(a.scale, tval[slot]);
                if (slot == dslot) v = add_rn(v, a.restart);
                tval[slot] = v;
                sq_add(pos, (double)v * (double)v);
                one = v != (T)0;
            }
            nz += __popc(__ballot_sync(kFull, one));
            if (a.hist) hist_values<T>(a.hist, win, one, v, pred_valid, pred0);
        }
        if (diag_new) {
            if (lane == 0) sq_add(pos, (double)a.restart * (double)a.restart);
            if (a.restart != (T)0) nz += 1;
            if (a.hist) hist_values<T>(a.hist, win, lane == 0 && a.restart != (T)0, a.restart, pred_valid, pred0);
        }
        __syncwarp();
        for (int idx = lane; idx < nk; idx += 32) {
            const int32_t j = kcol[idx];
            const T oldv = kval[idx];
            uint32_t slot = hash_slot(j, H);
            bool found = false;
            for (;;) {
                const int32_t kk = tkey[slot];
                if (kk == j) { found = true; break; }
                if (kk == -1) {
                    if (slot == (uint32_t)(HT - 1)) { slot = 0; continue; }
                    break;
                }
                ++slot;
            }
            T newv = (T)0;
            if (found) newv = tval[slot];
            else if (j == dcol) newv = a.restart;
            const T d = sub_rn(oldv, newv);
            sq_add(pos, (double)d * (double)d);
            sq_add(neg, (double)newv * (double)newv);
        }
        totals_add_rowsq(wt, sq_diff(sq_group_sum(pos, 32), sq_group_sum(neg, 32)));

        const int shift = diag_new ? 1 : 0;
        write_row(a, wt, row, base, n_cand, nz, [&](int i, int32_t& c, T& v) {
            if (i < shift) { c = dcol; v = a.restart; }
            else { const int slot = ord[i - shift]; c = tkey[slot]; v = tval[slot]; }
        });
        __syncwarp();
        for (int i = lane; i < count; i += 32) tkey[ord[i]] = -1;
        __syncwarp();
    }
    totals_flush(a, wt);
    __syncthreads();
    if (a.hist) hist_window_flush<T>(a.hist, win);
}

template <typename T>
__global__ void __launch_bounds__(256) propagate_kernel(PropArgs<T> a, PropShape sh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_wait();
    if (a.st->done) return;
    propagate_rows<T, false>(a, sh, smem_raw);
    pdl_launch_dependents();
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

// Rows deferred by the main kernel, one per warp at a time; gw = this warp's slot in the dense scratch.
// Ends with the warp's totals published; the caller flushes the histogram window after a CTA barrier.
template <typename T>
__device__ __forceinline__ void dense_rows(const PropArgs<T>& a, const DenseScratch<T>& sc, long long gw, unsigned int* win,
                                           bool win_mode, WinAcc& wacc, const WinStage& stage) {
    const unsigned n_fb = __ldcg(&a.st->fb_count);
    const int lane = lane_id();
    T* acc = sc.acc + gw * sc.n;
    T* oldv = sc.oldv + gw * sc.n;
    int32_t* ord = sc.ord + gw * sc.n;
    uint32_t* tbits = sc.tbits + gw * sc.words;
    uint32_t* kbits = sc.kbits + gw * sc.words;
    const bool cut_on = __ldcg(&a.st->cut_on) != 0;
    const T cut = (T)__ldcg(&a.st->cut);
    const bool pred_valid = a.prefill_l1 && __ldcg(&a.st->pred_valid) != 0;
    const typename KeyOf<T>::type pred0 = (typename KeyOf<T>::type)__ldcg(&a.st->pred_prefix0);
    const unsigned long long wlo = __ldcg(&a.st->win_lo), whi = __ldcg(&a.st->win_hi);
    WarpTotals wt;
    totals_init(wt);

    for (;;) {
        if (n_fb == 0) break;
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
        wt.kept += (unsigned long long)nk;
        wt.max_kept = max(wt.max_kept, nk);
        __syncwarp();

        const int32_t dcol = a.row_begin + row;
        const bool diag_new = ((tbits[dcol >> 5] >> (dcol & 31)) & 1u) == 0u;
        Sq128 pos = {0ull, 0ull}, neg = {0ull, 0ull};
        int nz = 0;
        for (int b = 0; b < count; b += 32) {
            int i = b + lane;
            bool one = false;
            T v = (T)0;
            if (i < count) {
                const int j = ord[i];
                v = mul_rn(a.scale, acc[j]);
                if (j == dcol) v = add_rn(v, a.restart);
                acc[j] = v;
                sq_add(pos, (double)v * (double)v);
                one = v != (T)0;
            }
            nz += __popc(__ballot_sync(kFull, one));
            if (win_mode) win_classify<T>(one, v, wlo, whi, wacc, stage, a.cand, a.cand_cap);
            else if (a.hist) hist_values<T>(a.hist, win, one, v, pred_valid, pred0);
        }
        if (diag_new) {
            if (lane == 0) sq_add(pos, (double)a.restart * (double)a.restart);
            if (a.restart != (T)0) nz += 1;
            if (win_mode) win_classify<T>(lane == 0 && a.restart != (T)0, a.restart, wlo, whi, wacc, stage, a.cand, a.cand_cap);
            else if (a.hist) hist_values<T>(a.hist, win, lane == 0 && a.restart != (T)0, a.restart, pred_valid, pred0);
        }
        __syncwarp();
        // old entries: every kept column has its bit in kbits and its value in oldv
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
                    sq_add(pos, (double)d * (double)d);
            sq_add(neg, (double)newv * (double)newv);
                }
            }
        }
        totals_add_rowsq(wt, sq_diff(sq_group_sum(pos, 32), sq_group_sum(neg, 32)));

        const int n_cand = count + (diag_new ? 1 : 0);
        const int shift = diag_new ? 1 : 0;
        const long long base = reserve_row(a, wt, n_cand);
        write_row(a, wt, row, base, n_cand, nz, [&](int i, int32_t& c, T& v) {
            if (i < shift) { c = dcol; v = a.restart; }
            else { c = ord[i - shift]; v = acc[c]; }
        });
        __syncwarp();
        // clear the bitmaps through the lists that set them
        for (int i = lane; i < count; i += 32) tbits[ord[i] >> 5] = 0u;
        for (int b = lane; b < len; b += 32) kbits[a.in.cols[ptr + b] >> 5] = 0u;
        // an operand read from a fixed slot (structure.cuh) belongs to a row that has just lost its structure: the slot
        // is abandoned and must not keep values for a later whole-array select pass
        if (ptr >= a.fixed_base) for (int b = lane; b < len; b += 32) a.in.vals[ptr + b] = (T)0;
        __syncwarp();
    }
    totals_flush(a, wt);
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) propagate_dense_kernel(PropArgs<T> a, DenseScratch<T> sc) {
    __shared__ unsigned int win[kHistWin];
    pdl_wait();
    if (a.st->done) return;
    // deferred rows can take long: the next kernel's CTAs are let in early only when there are none (CTAs waiting
    // in pdl_wait on the same SMs were measured to slow the rows down)
    if (a.st->fb_count == 0) pdl_launch_dependents();
    if (a.st->fb_count == 0) {
        // the usual case: no row was deferred, the step's histogram is already complete.  No CTA has row
        // work, so no "last CTA" is needed: CTA 0 scans, the others leave at once.
        if (a.fused_select && blockIdx.x == 0) scan_tail<T, WARPS * 32>(a.st, a.hist);
        return;
    }
    __shared__ unsigned int s_wn;
    for (int i = threadIdx.x; i < kHistWin; i += WARPS * 32) win[i] = 0u;
    if (threadIdx.x == 0) s_wn = 0u;
    __syncthreads();
    const bool win_mode = a.win_mode && a.st->win_on != 0;
    const WinStage stage = {reinterpret_cast<unsigned long long*>(win), &s_wn};
    WinAcc wacc = {0ull, ~0ull};
    dense_rows<T>(a, sc, (long long)blockIdx.x * WARPS + (threadIdx.x >> 5), win, win_mode, wacc, stage);
    __syncthreads();
    if (win_mode) win_flush(wacc, stage, a.cand, a.cand_cap, a.st);
    else if (a.hist) hist_window_flush<T>(a.hist, win);

    // ---- fused select (unsharded walk) when rows were deferred: the histogram is complete only now
    if (a.fused_select && !a.st->scan_done && last_block(&a.st->tick_dense)) scan_tail<T, WARPS * 32>(a.st, a.hist);
}

// Walk state for step 0 (utility.py:275-280), exchange buffer cleared, candidate block header reset.
__global__ void init_state_kernel(WalkState* st, double rp, double perc, double tol, int max_steps,
                                  unsigned long long* xbuf, int xbuf_len, unsigned long long* cand_block) {
    for (int i = threadIdx.x; i < xbuf_len; i += blockDim.x) xbuf[i] = 0ull;
    if (threadIdx.x == 0) {
        WalkState s;
        memset(&s, 0, sizeof(s));
        s.rp = rp; s.perc = perc; s.tol = tol; s.max_steps = max_steps;
        s.pool_gen = 1u;
        *st = s;
        cand_block[0] = 0ull;
        cand_block[1] = ~0ull;
    }
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
