// K2, symbolic / numeric form: the walk step's propagation (utility.py:282-283 fused with the lazy cut of
// utility.py:290) split the way sparse products are split when the same pattern is multiplied again and again.
//
// From the third walk step on the kept set of a restart row barely moves (the walk plateaus),
// yet a hash-accumulating kernel redoes the whole symbolic part of the row's SpGEMM every step.  Here the
// symbolic result is computed ONCE per row (symbolic_kernel) for a SUPERSET K+ of the kept columns (the row's
// entries above alpha x cut, alpha < 1) and cached in a pool:
//     O             the output columns K+ reaches, plus K+ itself and the restart column, sorted by the number
//                   of products that feed them (descending) - this is also the row's STORAGE ORDER from now on
//     streams       the outputs are dealt to G groups (G = 1, or 8 for a long row) and inside a group to the 32 lanes
//                   of a warp, chunk of 32 by chunk, in alternating direction (so every lane gets a similar number
//                   of products although the counts are heavily skewed); a lane's products - operand position in O,
//                   P value - follow each other output by output, inside an output in ascending operand column =
//                   exactly scipy's summation order, the last product of an output flagged.  Round j of a group's
//                   stream holds the j-th product of every lane that has one: contiguous, read once, front to back.
// and the row gets a FIXED slot of |O| entries in both iterate buffers.  A later step may use the structure
// whenever its kept set is INSIDE K+ (one bit per output): members of K+ at or below the cut take q = 0, their
// products are exact zeros and adding them changes no bit; outputs reached only through them come out 0 and
// count as absent, like csr_matmat drops them.  The numeric pass (numeric_stream) is then output-stationary: a
// lane adds the products of its outputs in order (separate multiply and add, as everywhere) - no hash table, no
// atomics, no sorting, no allocation, and the previous value of the same output (for ||Q - Q~||) sits at the same
// position of the input row.
//
// fast_kernel       short rows with a current structure, a warp each: the row's previous values and streams are
//                   brought into shared memory by bulk copies (TMA), two rows in flight per warp
// fast_big_kernel   long rows with a current structure, a CTA each (a warp per group), streams read from global memory
// symbolic_kernel   rows whose kept set left K+ (all rows on the first structure step): build the structure,
//                   numeric pass on the way
// Rows that exceed the symbolic kernel's tables go to the large-row kernel (fb_rows) and keep no structure.
#pragma once
#include "propagate.cuh"

namespace crwr {

constexpr int kFastWarps = 4;               // warps per CTA of the fast kernels
constexpr int kBigGroups = 8;               // groups of a long row (= warps of the CTA that replays it)
constexpr int kSymMaxCtas = 148 * 8;        // CTAs of one symbolic_kernel launch (each owns a global staging region)
constexpr unsigned kLastFlag = 0x8000u;     // stream entry: last product of its output

// Pool block of one row, nb = ceil(no / 32), G groups.  The part the numeric pass reads comes first, 16-byte aligned
// and a multiple of 16 bytes long, so that one bulk copy (TMA) brings it into shared memory:
//   [lload G*32 x u16][gptr (G+1) x i32][kbits nb x u32][ck nprod x u16](pad 16)[cw nprod x T](pad 16) | [ocols no x i32]
struct StructLayout { uint32_t lload, gptr, kbits, ck, cw, hot_bytes, ocols, bytes; };
template <typename T>
__host__ __device__ inline StructLayout struct_layout(int no, int nprod, int G) {
    const uint32_t nb = (uint32_t)((no + 31) >> 5);
    StructLayout L;
    uint32_t o = 0;
    L.lload = o; o += 2u * 32u * (uint32_t)G;
    L.gptr = o;  o += 4u * (uint32_t)(G + 1);
    L.kbits = o; o += 4u * nb;
    L.ck = o;    o += 2u * (uint32_t)nprod;
    o = (o + 15u) & ~15u;
    L.cw = o;    o += (uint32_t)sizeof(T) * (uint32_t)nprod;
    o = (o + 15u) & ~15u;
    L.hot_bytes = o;
    L.ocols = o; o += 4u * (uint32_t)no;
    L.bytes = (o + 15u) & ~15u;
    return L;
}
// The chunks of 32 outputs are dealt to the groups round by round in alternating direction, like the outputs of a chunk
// to the lanes: the counts fall from chunk to chunk, so a plain round robin would give group 0 the heaviest chunk of
// every round.  ii-th chunk of group g:
__device__ __forceinline__ int chunk_of(int ii, int g, int G) { return ii * G + ((ii & 1) ? G - 1 - g : g); }
// Position of the output that lane `lane` of group g owns in the group's ii-th chunk (the direction alternates).
__device__ __forceinline__ int lane_pos(int ii, int g, int G, int lane) {
    return (chunk_of(ii, g, G) << 5) + ((ii & 1) ? 31 - lane : lane);
}

// ---- bulk asynchronous copies global -> shared (TMA, cp.async.bulk) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Where a step's values go for the percentile select: histogram (early steps, sharded walks) or window (propagate.cuh,
// win_classify - restated here with keys of the data's own width and a 32-bit counter per lane).
template <typename T>
struct SelCtx {
    typedef typename KeyOf<T>::type K;
    bool win_mode, pred_valid;
    K klo, khi, succ;
    unsigned int below;
    K pred0;
    unsigned long long* hist;
    unsigned int* win;
    WinStage stage;
    unsigned long long* cand;
    unsigned long long cand_cap;
    __device__ __forceinline__ void add(bool nzv, T v) {          // all 32 lanes of the warp
        if (win_mode) {
            if (nzv) {
                const K key = KeyOf<T>::enc(v);
                if (key < klo) below += 1u;
                else if (key >= khi) succ = key < succ ? key : succ;
                else {
                    const unsigned int pos = atomicAdd(stage.count, 1u);
                    if (pos < (unsigned)kWinStage) stage.keys[pos] = (unsigned long long)key;
                    else {
                        const unsigned long long idx = atomicAdd(&cand[0], 1ull);
                        if (idx < cand_cap) cand[kCandHeader + idx] = (unsigned long long)key;
                    }
                }
            }
        } else if (hist) hist_values<T>(hist, win, nzv, v, pred_valid, pred0);
    }
    __device__ __forceinline__ void flush(WalkState* st) {        // CTA-wide
        if (win_mode) {
            WinAcc w;
            w.below = (unsigned long long)below;
            w.succ = succ == ~(K)0 ? ~0ull : (unsigned long long)succ;
            win_flush(w, stage, cand, cand_cap, st);
        } else if (hist) hist_window_flush<T>(hist, win);
    }
};
template <typename T>
__device__ __forceinline__ SelCtx<T> make_sel(const PropArgs<T>& a, unsigned int* win, unsigned int* win_count) {
    typedef typename KeyOf<T>::type K;
    SelCtx<T> s;
    WalkState* st = a.st;
    s.win_mode = a.win_mode && __ldcg(&st->win_on) != 0;
    s.klo = (K)__ldcg(&st->win_lo);
    s.khi = (K)__ldcg(&st->win_hi);
    s.succ = ~(K)0;
    s.below = 0u;
    s.pred_valid = a.prefill_l1 && __ldcg(&st->pred_valid) != 0;
    s.pred0 = (K)__ldcg(&st->pred_prefix0);
    s.hist = a.hist;
    s.win = win;
    s.stage.keys = reinterpret_cast<unsigned long long*>(win);
    s.stage.count = win_count;
    s.cand = a.cand;
    s.cand_cap = a.cand_cap;
    return s;
}

// What the rows of a thread add up to: the squares per lane (exact, so the grouping is free), the counts in lane 0.
// Reduced over the warp and published once per CTA.
struct RowTotals {
    Sq128 sq;
    unsigned long long nnz, kept;
    int max_kept, max_len;
};
__device__ __forceinline__ void publish_totals(WalkState* st, const RowTotals& t) {      // CTA-wide
    __shared__ unsigned long long s_tot[5];
    __shared__ int s_max[2];
    if (threadIdx.x < 5) s_tot[threadIdx.x] = 0ull;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();
    const Sq128 wsq = sq_group_sum(t.sq, 32);
    if (lane_id() == 0) {
        const unsigned long long sq0 = wsq.lo & 0xffffffffull, sq1 = wsq.lo >> 32, sq2 = wsq.hi;
        if (sq0) atomicAdd(&s_tot[0], sq0);
        if (sq1) atomicAdd(&s_tot[1], sq1);
        if (sq2) atomicAdd(&s_tot[2], sq2);
        if (t.nnz) atomicAdd(&s_tot[3], t.nnz);
        if (t.kept) atomicAdd(&s_tot[4], t.kept);
        if (t.max_kept) atomicMax(&s_max[0], t.max_kept);
        if (t.max_len) atomicMax(&s_max[1], t.max_len);
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
}

// ---------------------------------------------------------------------------------------------
// Numeric pass of one group of a row by one warp.  q_s (shared memory) holds the row's previous values in structure
// order with everything at or below the cut replaced by 0; res_s (shared memory) receives the sums.  The lane walks
// its share of the group's stream: round j of the stream holds the j-th product of every lane with more than j
// products, in lane order.  The loads of RN rounds are issued before the first of them is consumed (the sums
// themselves stay strictly in order).  Then the epilogue over the group's chunks: scale, restart, norm, select, store.
// SRC: kStaged - the structure lies in shared memory; kGlobalRO - in global memory, written by an earlier kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kStaged = 0, kGlobalRO = 1;

template <typename T, int SRC>
__device__ __forceinline__ void numeric_stream(const PropArgs<T>& a, const unsigned char* blk, int no, int nprod, int G, int g,
                                               int dpos, const T* q_s, T* res_s, long long slot, SelCtx<T>& sel, RowTotals& tot) {
    constexpr bool NC = SRC == kGlobalRO;
    constexpr int RN = SRC == kStaged ? 4 : 8;
    const StructLayout L = struct_layout<T>(no, nprod, G);
    const uint16_t* lload = reinterpret_cast<const uint16_t*>(blk + L.lload);
    const int32_t* gptr = reinterpret_cast<const int32_t*>(blk + L.gptr);
    const uint16_t* ck = reinterpret_cast<const uint16_t*>(blk + L.ck);
    const T* cw = reinterpret_cast<const T*>(blk + L.cw);
    const int lane = lane_id();
    const int nb = (no + 31) >> 5;
    const int load = NC ? (int)__ldg(lload + g * 32 + lane) : (int)lload[g * 32 + lane];
    int off = NC ? __ldg(gptr + g) : gptr[g];
    const int rounds = __reduce_max_sync(kFull, load);
    // The sums are parked by (chunk of the group, owning lane): res_s[ii * 32 * G + g * 32 + lane] - the group's ii-th
    // chunk occupies the 32 entries of chunk chunk_of(ii, g, G).  Outputs without a product keep 0.
    for (int ii = 0, c; (c = chunk_of(ii, g, G)) < nb; ++ii) res_s[(c << 5) + lane] = (T)0;
    __syncwarp();
    T acc = (T)0;
    T* res_cur = res_s + (g << 5) + lane;
    // (from an even chunk of the group to the next odd one and back: chunk_of)
    int res_step = (2 * G - 1 - 2 * g) << 5;
    const int res_flip = ((2 * G - 1 - 2 * g) << 5) ^ ((2 * g + 1) << 5);
    for (int j0 = 0; j0 < rounds; j0 += RN) {
        unsigned e[RN];
        T w[RN];
#pragma unroll
        for (int r = 0; r < RN; ++r) {
            const bool on = j0 + r < load;
            const unsigned m = __ballot_sync(kFull, on);
            const int idx = off + __popc(m & lanemask_lt());
            e[r] = 0u; w[r] = (T)0;
            if (on) { e[r] = NC ? __ldg(ck + idx) : ck[idx]; w[r] = NC ? __ldg(cw + idx) : cw[idx]; }
            off += __popc(m);
        }
#pragma unroll
        for (int r = 0; r < RN; ++r) {
            if (j0 + r < load) {
                acc = add_rn(acc, mul_rn(q_s[e[r] & (kLastFlag - 1u)], w[r]));
                if (e[r] & kLastFlag) { *res_cur = acc; res_cur += res_step; res_step ^= res_flip; acc = (T)0; }
            }
        }
    }
    __syncwarp();
    T* ov = a.out.vals + slot;
    Sq128 sq = tot.sq;
    int nz = 0;
    int ii = 0;
    for (int c; (c = chunk_of(ii, g, G)) < nb; ++ii) {
        const int p = (c << 5) + lane;
        const bool act = p < no;
        T v = (T)0;
        if (act) {
            v = mul_rn(a.scale, res_s[(c << 5) + ((ii & 1) ? 31 - lane : lane)]);      // the lane that owned this position
            if (p == dpos) v = add_rn(v, a.restart);
            const T d = sub_rn(q_s[p], v);
            sq_add(sq, (double)d * (double)d);
            ov[p] = v;
        }
        const bool nzv = act && v != (T)0;
        sel.add(nzv, v);
        nz += __popc(__ballot_sync(kFull, nzv));
    }
    tot.sq = sq;
    if (lane == 0) tot.nnz += (unsigned long long)nz;
}

// ---------------------------------------------------------------------------------------------
// The fast kernel: one warp per short row (one group), rows dealt round robin over a grid that fills the GPU once.
// Every warp runs a two-stage pipeline of its own: while row i is validated and summed out of shared memory, the bulk
// copies (TMA) of row i + stride - its previous values and the hot part of its pool block - are in flight, completing
// on the stage's mbarrier.  Long rows (several groups, or too large for a stage) belong to fast_big_kernel, which the
// host queues in front of this kernel once it knows of such rows; until then they are replayed here from global memory.
// ---------------------------------------------------------------------------------------------
constexpr int kFastStages = 2;
struct FastShape {
    int no_max;         // outputs per row a stage holds
    int hot_max;        // bytes of the pool block's hot part a stage holds (multiple of 16)
    size_t stage;       // bytes per stage
    size_t smem;        // dynamic shared memory per CTA
};
template <typename T>
__host__ __device__ inline FastShape fast_shape(int no_max, int hot_max) {
    FastShape f;
    f.no_max = no_max;
    f.hot_max = (hot_max + 15) & ~15;
    f.stage = align16((size_t)no_max * sizeof(T)) + (size_t)f.hot_max;
    // per warp: the stages and one array of sums
    f.smem = (size_t)kFastWarps * (kFastStages * f.stage + align16((size_t)no_max * sizeof(T))) + align16((size_t)kHistWin * 4);
    return f;
}
// Shared memory of fast_big_kernel: previous values and sums of one row of up to no_max outputs, the select window.
template <typename T>
__host__ __device__ inline size_t big_smem_bytes(int no_max) {
    return 2 * align16((size_t)no_max * sizeof(T)) + align16((size_t)kHistWin * 4);
}

// Validation of a row against its structure: the previous values (thresholded into q_s) must all lie inside K+.
// `lanes`: the threads sharing the row (a warp or the CTA), `t` this thread's index among them.
template <typename T>
__device__ __forceinline__ bool stage_previous(const T* rv, const uint32_t* kbits, int no, T cut, T* q_s, int t, int lanes, int& nkept) {
    bool bad = false;
    for (int b = 0; b < no; b += 4 * lanes) {
        T v[4];
        unsigned kb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = b + u * lanes + t;
            v[u] = (T)0; kb[u] = 0u;
            if (p < no) { v[u] = rv[p]; kb[u] = kbits[p >> 5]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = b + u * lanes + t;
            const bool kept = v[u] > cut;
            if (kept && ((kb[u] >> (p & 31)) & 1u) == 0u) bad = true;
            if (p < no) q_s[p] = kept ? v[u] : (T)0;
            nkept += __popc(__ballot_sync(kFull, kept));
        }
    }
    return bad;
}

template <typename T>
__global__ void __launch_bounds__(kFastWarps * 32) propagate_fast_kernel(PropArgs<T> a, FastShape fs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int s_wn;
    __shared__ __align__(8) unsigned long long s_bar[kFastWarps][kFastStages];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const size_t vals_bytes = align16((size_t)fs.no_max * sizeof(T));
    const size_t per_warp = kFastStages * fs.stage + vals_bytes;
    unsigned char* my = smem_raw + (size_t)warp * per_warp;
    T* res_s = reinterpret_cast<T*>(my + kFastStages * fs.stage);
    unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw + (size_t)kFastWarps * per_warp);
    for (int i = threadIdx.x; i < kHistWin; i += blockDim.x) win[i] = 0u;
    if (threadIdx.x == 0) s_wn = 0u;
    if (lane == 0) {
        for (int g = 0; g < kFastStages; ++g) mbar_init(&s_bar[warp][g], 1u);
        fence_mbar_init();
    }
    pdl_wait();
    WalkState* st = a.st;
    if (st->done) return;
    __syncthreads();
    const bool cut_on = __ldcg(&st->cut_on) != 0;
    const T cut = (T)__ldcg(&st->cut);
    const unsigned gen = __ldcg(&st->pool_gen);
    const int step = __ldcg(&st->step);
    SelCtx<T> sel = make_sel<T>(a, win, &s_wn);
    RowTotals tot = {{0ull, 0ull}, 0ull, 0ull, 0, 0};
    int n_fast = 0, n_big = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->unordered_rows = 1;

    // what the pipeline keeps of a header; kind 0: no usable structure (slow list), 1: staged through shared memory,
    // 2: replayed from global memory (a long row the big-row kernel has not taken), 3: already done this step
    struct Hdr { long long off, slot; int no, nprod, dpos, nk, hot, groups, kind; };
    auto load_hdr = [&](int row) -> Hdr {
        const RowStruct r = a.rs[row];
        Hdr h;
        h.off = r.off; h.slot = r.slot; h.no = r.no; h.nprod = r.nprod; h.dpos = r.dpos; h.nk = r.nk; h.hot = r.hot_bytes;
        h.groups = r.groups;
        h.kind = 0;
        if (cut_on && r.gen == gen && gen != 0u && r.no > 0) {
            if (r.step == step) h.kind = 3;
            else if (r.step == step - 1) h.kind = (r.groups == 1 && r.no <= fs.no_max && r.hot_bytes <= fs.hot_max) ? 1 : 2;
        }
        return h;
    };
    auto issue = [&](const Hdr& h, int g) {               // lane 0: the stage is free and behind a proxy fence
        const unsigned vb = (unsigned)align16((size_t)h.no * sizeof(T));
        const unsigned hb = (unsigned)h.hot;
        unsigned char* dst = my + (size_t)g * fs.stage;
        mbar_expect_tx(&s_bar[warp][g], vb + hb);
        bulk_g2s(dst, a.in.vals + h.slot, vb, &s_bar[warp][g]);
        bulk_g2s(dst + vals_bytes, a.pool + h.off, hb, &s_bar[warp][g]);
    };

    const int stride = gridDim.x * kFastWarps;
    const int row0 = blockIdx.x * kFastWarps + warp;
    // three headers in registers: the row at hand, the next (its copies are in flight), the one after (copies go out
    // when the row at hand has released its stage)
    Hdr h0, h1, h2;
    h0.kind = 0; h1.kind = 0; h2.kind = 0;
    if (row0 < a.n_rows) h0 = load_hdr(row0);
    if (row0 + stride < a.n_rows) h1 = load_hdr(row0 + stride);
    if (row0 + 2 * stride < a.n_rows) h2 = load_hdr(row0 + 2 * stride);
    if (h0.kind == 1 && lane == 0) issue(h0, 0);
    if (h1.kind == 1 && lane == 0) issue(h1, 1);
    unsigned phase = 0u;        // bit g: parity the next wait on stage g expects
    int g = 0;
    for (int row = row0; row < a.n_rows; row += stride) {
        const Hdr h = h0;
        const int kind = h.kind;
        unsigned char* stg = my + (size_t)g * fs.stage;
        T* q_s = reinterpret_cast<T*>(stg);
        bool ok = kind == 1 || kind == 2;
        int nkept = 0;
        if (kind == 1) {
            mbar_wait(&s_bar[warp][g], (phase >> g) & 1u);
            phase ^= 1u << g;
        }
        if (ok) {
            const StructLayout L = struct_layout<T>(h.no, h.nprod, h.groups);
            const uint32_t* kbits = kind == 1 ? reinterpret_cast<const uint32_t*>(stg + vals_bytes + L.kbits)
                                              : reinterpret_cast<const uint32_t*>(a.pool + h.off + L.kbits);
            const T* rv = kind == 1 ? q_s : a.in.vals + h.slot;
            bool bad = true;                                        // no room for the previous values: rebuild
            if (h.no <= fs.no_max) bad = stage_previous<T>(rv, kbits, h.no, cut, q_s, lane, 32, nkept);
            ok = !__any_sync(kFull, bad);
        }
        if (ok) {
            __syncwarp();
            if (kind == 1) numeric_stream<T, kStaged>(a, stg + vals_bytes, h.no, h.nprod, 1, 0, h.dpos, q_s, res_s, h.slot, sel, tot);
            else {
                for (int gg = 0; gg < h.groups; ++gg)
                    numeric_stream<T, kGlobalRO>(a, a.pool + h.off, h.no, h.nprod, h.groups, gg, h.dpos, q_s, res_s, h.slot, sel, tot);
                ++n_big;
            }
            if (lane == 0) {
                a.out.row_ptr[row] = h.slot;
                a.out.row_len[row] = h.no;
                a.rs[row].step = step;
                tot.kept += (unsigned long long)nkept;
                tot.max_len = max(tot.max_len, h.no);
                tot.max_kept = max(tot.max_kept, h.nk);
            }
            ++n_fast;
        } else if (kind != 3 && lane == 0) {
            a.slow_list[atomicAdd(&st->slow_count, 1u)] = row;
        }
        // the stage is released: copies of the row after next, header of the one after that
        fence_proxy_async();
        __syncwarp();
        h0 = h1;
        h1 = h2;
        if (row + 2 * stride >= a.n_rows) h1.kind = 0;
        if (h1.kind == 1 && lane == 0) issue(h1, g);
        h2.kind = 0;
        if (row + 3 * stride < a.n_rows) h2 = load_hdr(row + 3 * stride);
        g ^= 1;
    }
    pdl_launch_dependents();
    if (lane == 0 && n_fast) atomicAdd(&st->fast_rows, (unsigned)n_fast);
    if (lane == 0 && n_big) atomicAdd(&st->big_rows, (unsigned)n_big);
    publish_totals(st, tot);
    sel.flush(st);
}

// ---------------------------------------------------------------------------------------------
// Long rows: one CTA per row, one warp per group of the row.  The previous values are validated and thresholded into
// shared memory by the whole CTA; every warp then replays its group's stream from global memory (read once, front to
// back, kNumR rounds of loads in flight).  Rows this kernel takes are marked done for the step; the fast kernel that
// follows skips them.  fs: the fast kernel's stage limits (a row that fits them is left to it).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBigGroups * 32, 4) propagate_fast_big_kernel(PropArgs<T> a, FastShape fs, int NQ) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int s_wn;
    __shared__ int s_nkept;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = lane_id();
    const int warp = tid >> 5;
    const int W = nth >> 5;
    T* q_s = reinterpret_cast<T*>(smem_raw);
    T* res_s = reinterpret_cast<T*>(smem_raw + align16((size_t)NQ * sizeof(T)));
    unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw + 2 * align16((size_t)NQ * sizeof(T)));
    for (int i = tid; i < kHistWin; i += nth) win[i] = 0u;
    if (tid == 0) s_wn = 0u;
    pdl_wait();
    WalkState* st = a.st;
    if (st->done) return;
    __syncthreads();
    const bool cut_on = __ldcg(&st->cut_on) != 0;
    const T cut = (T)__ldcg(&st->cut);
    const unsigned gen = __ldcg(&st->pool_gen);
    const int step = __ldcg(&st->step);
    SelCtx<T> sel = make_sel<T>(a, win, &s_wn);
    RowTotals tot = {{0ull, 0ull}, 0ull, 0ull, 0, 0};
    int n_done = 0;
    if (cut_on && gen != 0u)
    for (int row = blockIdx.x; row < a.n_rows; row += gridDim.x) {
        const RowStruct h = a.rs[row];
        const bool mine = h.gen == gen && h.no > 0 && h.no <= NQ && h.step == step - 1 &&
                          !(h.groups == 1 && h.no <= fs.no_max && h.hot_bytes <= fs.hot_max);
        if (!mine) continue;                                    // (uniform over the CTA)
        if (tid == 0) s_nkept = 0;
        __syncthreads();
        const StructLayout L = struct_layout<T>(h.no, h.nprod, h.groups);
        int nkept = 0;
        const bool bad = stage_previous<T>(a.in.vals + h.slot, reinterpret_cast<const uint32_t*>(a.pool + h.off + L.kbits), h.no, cut,
                                           q_s, tid, nth, nkept);
        if (lane == 0 && nkept) atomicAdd(&s_nkept, nkept);
        if (__syncthreads_or(bad ? 1 : 0)) {
            // the kept set left K+: the fast kernel finds the row neither current nor done and hands it to the symbolic kernel
            if (tid == 0) a.rs[row].step = -2;
            continue;
        }
        for (int g = warp; g < h.groups; g += W)
            numeric_stream<T, kGlobalRO>(a, a.pool + h.off, h.no, h.nprod, h.groups, g, h.dpos, q_s, res_s, h.slot, sel, tot);
        if (tid == 0) {
            a.out.row_ptr[row] = h.slot;
            a.out.row_len[row] = h.no;
            a.rs[row].step = step;
            tot.kept += (unsigned long long)s_nkept;
            tot.max_len = max(tot.max_len, h.no);
            tot.max_kept = max(tot.max_kept, h.nk);
            ++n_done;
        }
        __syncthreads();
    }
    pdl_launch_dependents();
    if (tid == 0 && n_done) { atomicAdd(&st->fast_rows, (unsigned)n_done); atomicAdd(&st->big_rows, (unsigned)n_done); }
    publish_totals(st, tot);
    sel.flush(st);
}

// ---------------------------------------------------------------------------------------------
// The symbolic kernel: one CTA per row of the slow list (every row on the first structure step).  Everything is
// parallel over the row's PRODUCTS (operand k of K+ times entry of P row k), which are dealt to the threads in stream
// order (k ascending, P-row order):
//   P1  K+ = the row's entries above alpha x cut (q = the value where it is above the cut, else 0), sorted by column
//   P2  hash table of the output set (K+ itself, the restart column, every column a product reaches) with the number
//       of products per output; slot and operand of every product are remembered
//   P3  outputs sorted by count, descending (counting sort); block offsets
//   P4  the products into one staging list per output, IN STREAM ORDER: the operands are taken in windows of kWinOps; the
//       products of a window set their operand's bit in the output's window mask, a product's place is the output's
//       fill level before the window plus the number of lower bits in the mask (one mask bit per operand of the window)
//   P5  pool block: the outputs dealt to groups and lanes, the lanes' products written as the groups' streams (the
//       writer walks the rounds exactly as the reader will); fixed slot; header
//   P6  numeric pass, fused into P5: the lane that copies a product also adds it
// Nothing depends on the order in which threads run, except the order of the outputs inside a count class (arbitrary;
// it only decides where a value is stored, never its bits).
// ---------------------------------------------------------------------------------------------
struct SymShape {
    int H;              // hash slots (power of two, >= 2 * no_max + 64)
    int no_max;         // outputs a row may have
    int km;             // operand columns (|K+|) a row may have
    int threads;        // CTA size for short lists (the first build uses threads_build)
    int threads_build;
    int tmp_smem;       // products staged in shared memory; longer rows stage in global memory
    int tmp_cap;        // products per global staging region (one per CTA)
    size_t smem;        // dynamic shared memory per CTA (window included)
};
// Operands per ordering window of P4 = bits of an output's window mask.  (64-bit masks halve the windows of a long row but
// measured no faster, and cost 4 bytes per hash slot of shared memory that long rows need for their tables.)
typedef unsigned int WMask;
constexpr int kWinOps = 8 * (int)sizeof(WMask);
__device__ __forceinline__ int wmask_popc(unsigned int m) { return __popc(m); }
__device__ __forceinline__ int wmask_popc(unsigned long long m) { return __popcll(m); }

__host__ __device__ inline size_t sym_smem_bytes(int H, int no_max, int km, int tmp_smem, size_t s) {
    const size_t nbm = (size_t)(no_max + 31) / 32;
    return align16((size_t)H * 4) * 2 + align16((size_t)H * sizeof(WMask)) + align16((size_t)H * 2) +     // hkey, hcnt, wmask, hpos
           align16((size_t)km * 4) * 2 + align16((size_t)km * s) + align16((size_t)(km + 2) * 4) * 2 +   // kcol, kps, kval, kpre, bins
           align16((size_t)km * 2) * 2 +                              // kslot, kpos
           align16((size_t)no_max * 2) * 3 +                          // olist, oslot, cnt_s
           align16((size_t)no_max * 4) +                              // cptr
           align16((size_t)no_max * s) * 2 +                          // q_s, res_s
           align16((nbm + 1) * 4) * 2 +                               // bptr_s, kb_s
           align16((size_t)tmp_smem * 2) * 3 + align16((size_t)tmp_smem * s) +   // pslot, pk, tck, tcw
           align16((size_t)kHistWin * 4);                             // select window
}
__host__ __device__ inline size_t sym_tmp_bytes(int tmp_cap, size_t s) {        // one CTA's global staging region
    return align16((size_t)tmp_cap * 2) * 3 + align16((size_t)tmp_cap * s);
}

__device__ __forceinline__ uint32_t sym_hash(int32_t j, int mask) { return (((uint32_t)j * 0x9E3779B1u) >> 7) & (uint32_t)mask; }

template <typename T>
__global__ void __launch_bounds__(512) symbolic_kernel(PropArgs<T> a, SymShape sh) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int s_wn;
    __shared__ int s_nk, s_nkt, s_nout, s_nprod;
    __shared__ uint16_t lload_s[32 * kBigGroups];
    __shared__ int32_t gptr_s[kBigGroups + 1];
    __shared__ unsigned long long s_slot_base, s_slot_left, s_pool_base, s_pool_left, s_slot_rel, s_poff;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = lane_id();
    const int warp = tid >> 5;
    const int W = nth >> 5;
    const int H = sh.H, HM = sh.H - 1, NOMAX = sh.no_max, KM = sh.km;
    const size_t s = sizeof(T);
    unsigned char* p = smem_raw;
    int32_t* hkey = reinterpret_cast<int32_t*>(p);      p += align16((size_t)H * 4);
    int32_t* hcnt = reinterpret_cast<int32_t*>(p);      p += align16((size_t)H * 4);
    WMask* wmask = reinterpret_cast<WMask*>(p);         p += align16((size_t)H * sizeof(WMask));
    uint16_t* hpos = reinterpret_cast<uint16_t*>(p);    p += align16((size_t)H * 2);
    int32_t* kcol = reinterpret_cast<int32_t*>(p);      p += align16((size_t)KM * 4);
    int32_t* kps = reinterpret_cast<int32_t*>(p);       p += align16((size_t)KM * 4);
    T* kval = reinterpret_cast<T*>(p);                  p += align16((size_t)KM * s);
    int32_t* kpre = reinterpret_cast<int32_t*>(p);      p += align16((size_t)(KM + 2) * 4);
    int32_t* bins = reinterpret_cast<int32_t*>(p);      p += align16((size_t)(KM + 2) * 4);
    uint16_t* kslot = reinterpret_cast<uint16_t*>(p);   p += align16((size_t)KM * 2);
    uint16_t* kpos = reinterpret_cast<uint16_t*>(p);    p += align16((size_t)KM * 2);
    uint16_t* olist = reinterpret_cast<uint16_t*>(p);   p += align16((size_t)NOMAX * 2);
    uint16_t* oslot = reinterpret_cast<uint16_t*>(p);   p += align16((size_t)NOMAX * 2);
    uint16_t* cnt_s = reinterpret_cast<uint16_t*>(p);   p += align16((size_t)NOMAX * 2);
    int32_t* cptr = reinterpret_cast<int32_t*>(p);      p += align16((size_t)NOMAX * 4);
    T* q_s = reinterpret_cast<T*>(p);                   p += align16((size_t)NOMAX * s);
    T* res_s = reinterpret_cast<T*>(p);                 p += align16((size_t)NOMAX * s);
    const size_t nbm = (size_t)(NOMAX + 31) / 32;
    int32_t* bptr_s = reinterpret_cast<int32_t*>(p);    p += align16((nbm + 1) * 4);
    uint32_t* kb_s = reinterpret_cast<uint32_t*>(p);    p += align16((nbm + 1) * 4);
    uint16_t* pslot_s = reinterpret_cast<uint16_t*>(p); p += align16((size_t)sh.tmp_smem * 2);
    uint16_t* pk_s = reinterpret_cast<uint16_t*>(p);    p += align16((size_t)sh.tmp_smem * 2);
    uint16_t* tck_s = reinterpret_cast<uint16_t*>(p);   p += align16((size_t)sh.tmp_smem * 2);
    T* tcw_s = reinterpret_cast<T*>(p);                 p += align16((size_t)sh.tmp_smem * s);
    unsigned int* win = reinterpret_cast<unsigned int*>(p);
    // the unsorted K+ is parked in cptr / q_s (KM <= NOMAX) until it has been sorted into kcol / kval
    int32_t* kcol_u = cptr;
    T* kval_u = q_s;

    for (int i = tid; i < kHistWin; i += nth) win[i] = 0u;
    for (int i = tid; i < H; i += nth) { hkey[i] = -1; hcnt[i] = 0; wmask[i] = (WMask)0; }
    if (tid == 0) { s_wn = 0u; s_slot_left = 0ull; s_pool_left = 0ull; s_slot_base = 0ull; s_pool_base = 0ull; }
    pdl_wait();
    WalkState* st = a.st;
    if (st->done) return;
    __syncthreads();
    const bool cut_on = __ldcg(&st->cut_on) != 0;
    const T cut = (T)__ldcg(&st->cut);
    const T cut_lo = (T)((double)a.alpha * (double)cut);
    const unsigned gen = __ldcg(&st->pool_gen);
    const int step = __ldcg(&st->step);
    const int n_list = a.row_list_count ? (int)__ldcg(a.row_list_count) : a.n_rows;
    SelCtx<T> sel = make_sel<T>(a, win, &s_wn);
    RowTotals tot = {{0ull, 0ull}, 0ull, 0ull, 0, 0};
    int n_built = 0, n_big = 0;
    long long hot_delta = 0;
    if (blockIdx.x == 0 && tid == 0) {
        st->unordered_rows = 1;
        if (!cut_on && n_list > 0) st->order_hazard = 1;      // structures only exist behind a cut (never for positive data)
    }
    // few rows (settled walk): exact reservations; many (first build): a CTA reserves for several rows at a time
    const bool chunked = n_list > 2 * (int)gridDim.x;
    unsigned char* gtmp = a.sym_tmp + (size_t)blockIdx.x * sym_tmp_bytes(sh.tmp_cap, s);
    uint16_t* pslot_g = reinterpret_cast<uint16_t*>(gtmp);              gtmp += align16((size_t)sh.tmp_cap * 2);
    uint16_t* pk_g = reinterpret_cast<uint16_t*>(gtmp);                 gtmp += align16((size_t)sh.tmp_cap * 2);
    uint16_t* tck_g = reinterpret_cast<uint16_t*>(gtmp);                gtmp += align16((size_t)sh.tmp_cap * 2);
    T* tcw_g = reinterpret_cast<T*>(gtmp);
    const int32_t* __restrict__ pcols = a.p_cols;
    const T* __restrict__ pvals = a.p_vals;

    auto find = [&](int32_t j) -> int {             // slot of a column that is in the table
        int sl = (int)sym_hash(j, HM);
        CRWR_GUARD_DECL(g);
        while (hkey[sl] != j) { sl = (sl + 1) & HM; CRWR_GUARD(st, g, H, 7, break); }
        return sl;
    };
    auto insert = [&](int32_t j) -> int {
        int sl = (int)sym_hash(j, HM);
        for (;;) {
            const int32_t k = hkey[sl];
            if (k == j) return sl;
            if (k == -1) {
                const int32_t o = atomicCAS(&hkey[sl], -1, j);
                if (o == -1) {
                    const int idx = atomicAdd(&s_nout, 1);
                    if (idx < NOMAX) olist[idx] = (uint16_t)sl;
                    return sl;
                }
                if (o == j) return sl;
            }
            sl = (sl + 1) & HM;
        }
    };
    // entries of a reserved chunk of the slot region that no row received: zero (a whole-array select pass covers them)
    auto zero_slots = [&](unsigned long long base, unsigned long long left) {
        for (unsigned long long i = tid; i < left && base + i < (unsigned long long)a.slot_cap; i += nth) {
            a.in.vals[a.fixed_base + (long long)(base + i)] = (T)0;
            a.out.vals[a.fixed_base + (long long)(base + i)] = (T)0;
        }
    };

    if (cut_on)
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int row = a.row_list_count ? __ldg(a.row_order + li) : li;
        const int32_t dcol = a.row_begin + row;
        const RowStruct old = a.rs[row];
        if (tid == 0) { s_nk = 0; s_nkt = 0; s_nout = 0; }
        __syncthreads();
        // ---- P1: K+ (unordered), then sorted by column (rank by counting)
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        for (int b = 0; b < len; b += nth) {
            const int e = b + tid;
            T v = (T)0;
            int32_t c = 0;
            if (e < len) { v = __ldg(a.in.vals + ptr + e); c = __ldg(a.in.cols + ptr + e); }
            const bool keep = v > cut_lo;
            const bool kt = v > cut;
            const unsigned m = __ballot_sync(kFull, keep);
            const unsigned mt = __ballot_sync(kFull, kt);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_nk, __popc(m));
            if (lane == 0 && mt) atomicAdd(&s_nkt, __popc(mt));
            base = __shfl_sync(kFull, base, 0);
            const int pos = base + __popc(m & lanemask_lt());
            if (keep && pos < KM) { kcol_u[pos] = c; kval_u[pos] = kt ? v : (T)0; }
        }
        // the row's old fixed slot (if any) is abandoned: no stale value may be left for a whole-array select pass.
        // The copy in this step's output buffer now; the one in the input buffer (the row's operand) once it is no
        // longer needed - below, or by the large-row kernel when the row is deferred to it.
        const bool had_slot = old.gen == gen && old.no > 0;
        if (had_slot) {
            for (int i = tid; i < old.no; i += nth) a.out.vals[old.slot + i] = (T)0;
            if (tid == 0) { a.rs[row].gen = 0u; hot_delta -= (long long)old.hot_bytes; }
        }
        __syncthreads();
        const int nk = s_nk, nkt = s_nkt;
        bool defer = nk > KM;
        int no = 0, nprod = 0;
        if (!defer) {
            // (kcol_u is padded to a multiple of four with a column no real one is below)
            for (int i = nk + tid; i < ((nk + 3) & ~3); i += nth) kcol_u[i] = 0x7fffffff;
            __syncthreads();
            for (int i = tid; i < nk; i += nth) {
                const int32_t c = kcol_u[i];
                int r = 0;
                const int4* q4 = reinterpret_cast<const int4*>(kcol_u);
                for (int j = 0; j < (nk + 3) >> 2; ++j) {
                    const int4 v = q4[j];
                    r += (v.x < c) + (v.y < c) + (v.z < c) + (v.w < c);
                }
                kcol[r] = c;
                kval[r] = kval_u[i];
                const int ps = __ldg(a.p_indptr + c);
                kps[r] = ps;
                kpre[r + 1] = __ldg(a.p_indptr + c + 1) - ps;         // lengths first, prefix sums next
            }
            for (int i = tid; i <= nk; i += nth) bins[i] = 0;
            __syncthreads();
            if (warp == 0) {
                int run = 0;
                for (int b = 0; b < nk; b += 32) {
                    const int i = b + lane;
                    const int l = i < nk ? kpre[i + 1] : 0;
                    int incl = l;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (i < nk) kpre[i + 1] = run + incl;
                    run += __shfl_sync(kFull, incl, 31);
                }
                if (lane == 0) { kpre[0] = 0; s_nprod = run; }
            }
            __syncthreads();
            nprod = s_nprod;
            if (nprod > sh.tmp_smem && nprod > sh.tmp_cap) defer = true;
        }
        uint16_t* pslot = nprod <= sh.tmp_smem ? pslot_s : pslot_g;
        uint16_t* pk = nprod <= sh.tmp_smem ? pk_s : pk_g;
        uint16_t* tck = nprod <= sh.tmp_smem ? tck_s : tck_g;
        T* tcw = nprod <= sh.tmp_smem ? tcw_s : tcw_g;
        if (!defer) {
            // ---- P2: the output set and the products per output
            for (int i = tid; i < nk; i += nth) kslot[i] = (uint16_t)insert(kcol[i]);
            if (tid == 0) insert(dcol);
            // a warp per operand: its P row is read coalesced; product i = kpre[k] + e of the stream.  The first 32 entries of
            // the warp's next operand are requested before the current operand's are inserted.
            {
                int k = warp;
                int32_t j_n = -1;
                if (k < nk && lane < kpre[k + 1] - kpre[k]) j_n = __ldg(pcols + kps[k] + lane);
                for (; k < nk; k += W) {
                    const int ps = kps[k], i0 = kpre[k], ln = kpre[k + 1] - i0;
                    int32_t j = j_n;
                    const int kn = k + W;
                    j_n = -1;
                    if (kn < nk && lane < kpre[kn + 1] - kpre[kn]) j_n = __ldg(pcols + kps[kn] + lane);
                    for (int e = lane; e < ln; e += 32) {
                        if (e >= 32) j = __ldg(pcols + ps + e);
                        int sl = 0;
                        if (s_nout <= NOMAX) {
                            CRWR_CHECK(st, j >= 0, 5);
                            sl = insert(j);
                            atomicAdd(&hcnt[sl], 1);
                        }
                        pslot[i0 + e] = (uint16_t)sl;
                        pk[i0 + e] = (uint16_t)k;
                    }
                }
            }
            __syncthreads();
            no = s_nout;
            if (no > NOMAX) defer = true;
        }
        if (!defer) {
            // ---- P3: counting sort by count, descending; block offsets
            for (int o = tid; o < no; o += nth) atomicAdd(&bins[hcnt[olist[o]]], 1);
            __syncthreads();
            if (warp == 0) {
                int run = 0;
                for (int hi = nk; hi >= 0; hi -= 32) {
                    const int c = hi - lane;
                    const int v = c >= 0 ? bins[c] : 0;
                    int incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += u;
                    }
                    if (c >= 0) bins[c] = run + incl - v;
                    run += __shfl_sync(kFull, incl, 31);
                }
            }
            __syncthreads();
            for (int o = tid; o < no; o += nth) {
                const int sl = olist[o];
                const int c = hcnt[sl];
                const int pos = atomicAdd(&bins[c], 1);
                oslot[pos] = (uint16_t)sl;
                cnt_s[pos] = (uint16_t)c;
                hpos[sl] = (uint16_t)pos;
                hcnt[sl] = 0;                   // from here on: the output's fill level
            }
            __syncthreads();
            const int nb = (no + 31) >> 5;
            for (int b = warp; b < nb; b += W) {
                const int pp = (b << 5) + lane;
                const int c = pp < no ? (int)cnt_s[pp] : 0;
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += u;
                }
                if (pp < no) { cptr[pp] = incl - c; q_s[pp] = (T)0; }     // relative to the block's offset
                if (lane == 31) { bptr_s[b] = incl; kb_s[b] = 0u; }
            }
            __syncthreads();
            if (warp == 0) {
                int run = 0;
                for (int b0 = 0; b0 <= nb; b0 += 32) {
                    const int b = b0 + lane;
                    const int v = b < nb ? bptr_s[b] : 0;
                    int incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += u;
                    }
                    if (b <= nb) bptr_s[b] = run + incl - v;
                    run += __shfl_sync(kFull, incl, 31);
                }
            }
            __syncthreads();
            CRWR_CHECK(st, bptr_s[nb] == nprod, 8);
        }
        unsigned long long slot_rel = 0, poff = 0;
        // a row the fast kernel cannot stage in shared memory is dealt to several groups (a CTA replays it)
        int G = 1;
        if (no > a.small_no || struct_layout<T>(no, nprod, 1).hot_bytes > (uint32_t)a.small_hot) G = kBigGroups;
        const StructLayout L = struct_layout<T>(no, nprod, G);
        const unsigned long long no4 = (unsigned long long)((no + 3) & ~3);     // slots start on 16-byte boundaries (bulk copies)
        // A rebuilt structure that fits the row's old block and slot stays in place (blocks and slots are never handed
        // back during a walk, and a slowly settling walk rebuilds thousands of rows in its first steps); new ones are
        // reserved with a quarter to spare for that.
        const bool reuse = !defer && had_slot && (unsigned long long)L.bytes <= (unsigned long long)old.blk_cap &&
                           no4 <= (unsigned long long)old.slot_cap;
        const unsigned long long want_slot = reuse ? (unsigned long long)old.slot_cap : ((no4 + no4 / 4 + 3ull) & ~3ull);
        const unsigned long long want_blk = reuse ? (unsigned long long)old.blk_cap
                                                  : (((unsigned long long)L.bytes + (unsigned long long)L.bytes / 4 + 15ull) & ~15ull);
        if (reuse) {
            slot_rel = (unsigned long long)(old.slot - a.fixed_base);
            poff = (unsigned long long)old.off;
        } else if (!defer) {
            // ---- reservations: fixed slot in both iterate buffers, pool block (thread 0; chunked while many rows are built)
            if (want_slot > s_slot_left) zero_slots(s_slot_base, s_slot_left);
            __syncthreads();
            if (tid == 0) {
                if (want_slot > s_slot_left) {
                    const unsigned long long sz = (chunked && (unsigned long long)a.slot_chunk > want_slot) ? (unsigned long long)a.slot_chunk : want_slot;
                    s_slot_base = atomicAdd(&st->slot_cursor, sz);
                    s_slot_left = sz;
                }
                s_slot_rel = s_slot_base; s_slot_base += want_slot; s_slot_left -= want_slot;
                if (want_blk > s_pool_left) {
                    const unsigned long long sz = (chunked && (unsigned long long)a.pool_chunk > want_blk) ? (unsigned long long)a.pool_chunk : want_blk;
                    s_pool_base = atomicAdd(&st->pool_cursor, sz);
                    s_pool_left = sz;
                }
                s_poff = s_pool_base; s_pool_base += want_blk; s_pool_left -= want_blk;
            }
            __syncthreads();
            slot_rel = s_slot_rel; poff = s_poff;
            if (slot_rel + want_slot > (unsigned long long)a.slot_cap || poff + want_blk > a.pool_cap) {
                // slot region or pool exhausted: the row goes without a structure (large-row kernel), this step and - as
                // the cursors only grow - every later one; struct_full tells the host a larger workspace would be faster
                if (tid == 0) st->struct_full = 1;
                zero_slots(slot_rel, want_slot);    // the slot entries it did get stay empty
                defer = true;
            }
        }
        if (defer) {
            // too large for the tables: the large-row kernel computes the row, it keeps no structure
            if (no > NOMAX) { for (int i = tid; i < H; i += nth) { hkey[i] = -1; hcnt[i] = 0; } }
            else for (int i = tid; i < no; i += nth) { const int sl = olist[i]; hkey[sl] = -1; hcnt[sl] = 0; }
            if (tid == 0) { a.fb_rows[atomicAdd(&st->fb_count, 1u)] = row; tot.max_kept = max(tot.max_kept, nk); }
            __syncthreads();
            continue;
        }
        const long long slot = a.fixed_base + (long long)slot_rel;
        unsigned char* blk = a.pool + poff;
        const int nb = (no + 31) >> 5;
        // (everything of the slot beyond the row's entries is zero in both buffers: a new slot's spare part here; a reused
        // slot's was cleared when it was reserved, and its old entries above and below)
        if (!reuse) for (int i = no + tid; i < (int)want_slot; i += nth) { a.in.vals[slot + i] = (T)0; a.out.vals[slot + i] = (T)0; }
        if (had_slot) for (int i = tid; i < old.no; i += nth) a.in.vals[old.slot + i] = (T)0;
        // positions of the operands, K+ membership bits, previous values in structure order
        for (int i = tid; i < nk; i += nth) {
            const int pos = hpos[kslot[i]];
            kpos[i] = (uint16_t)pos;
            q_s[pos] = kval[i];
            atomicOr(&kb_s[pos >> 5], 1u << (pos & 31));
        }
        const int dpos = hpos[find(dcol)];
        __syncthreads();
        // ---- P4: the products into the staging lists, in stream order, one window of 32 operands at a time
        for (int w0 = 0; w0 < nk; w0 += kWinOps) {
            const int i_lo = kpre[w0], i_hi = kpre[min(w0 + kWinOps, nk)];
            for (int i = i_lo + tid; i < i_hi; i += nth) atomicOr(&wmask[pslot[i]], (WMask)1 << ((int)pk[i] - w0));
            __syncthreads();
            for (int i = i_lo + tid; i < i_hi; i += nth) {
                const int sl = pslot[i], k = pk[i];
                const WMask bit = (WMask)1 << (k - w0);
                const int pos = hpos[sl];
                const int at = bptr_s[pos >> 5] + cptr[pos] + hcnt[sl] + wmask_popc(wmask[sl] & (WMask)(bit - 1));
                CRWR_CHECK(st, at >= 0 && at < nprod, 9);
                tck[at] = kpos[k];
                tcw[at] = __ldg(pvals + kps[k] + (i - kpre[k]));
            }
            __syncthreads();
            for (int i = i_lo + tid; i < i_hi; i += nth) {
                const int sl = pslot[i];
                const WMask m = wmask[sl];
                const WMask bit = (WMask)1 << ((int)pk[i] - w0);
                // the output's first product of the window updates its fill level (a later one may already see the mask cleared)
                if (m != 0 && (m & (WMask)(bit - 1)) == 0) { hcnt[sl] += wmask_popc(m); wmask[sl] = (WMask)0; }
            }
            __syncthreads();
        }
        // ---- P5: the lanes' shares (chunk by chunk, alternating direction), the group streams, the slot's columns
        // (written once, both buffers); P6, the numeric pass, on the way: the lane that copies a product also adds it
        int32_t* ocols_w = reinterpret_cast<int32_t*>(blk + L.ocols);
        uint16_t* lload_w = reinterpret_cast<uint16_t*>(blk + L.lload);
        int32_t* gptr_w = reinterpret_cast<int32_t*>(blk + L.gptr);
        uint32_t* kbits_w = reinterpret_cast<uint32_t*>(blk + L.kbits);
        uint16_t* ck_w = reinterpret_cast<uint16_t*>(blk + L.ck);
        T* cw_w = reinterpret_cast<T*>(blk + L.cw);
        for (int g = warp; g < G; g += W) {
            int sum = 0;
            for (int ii = 0; chunk_of(ii, g, G) < nb; ++ii) {
                const int pp = lane_pos(ii, g, G, lane);
                if (pp < no) sum += (int)cnt_s[pp];
            }
            lload_s[g * 32 + lane] = (uint16_t)sum;
            CRWR_CHECK(st, sum < 65536, 10);
            const int gsum = __reduce_add_sync(kFull, sum);
            if (lane == 0) gptr_s[g + 1] = gsum;
        }
        for (int pp = tid; pp < no; pp += nth) {
            const int32_t col = hkey[oslot[pp]];
            ocols_w[pp] = col;
            a.in.cols[slot + pp] = col;
            a.out.cols[slot + pp] = col;
            res_s[pp] = (T)0;
        }
        for (int i = tid; i < nb; i += nth) kbits_w[i] = kb_s[i];
        __syncthreads();
        if (tid == 0) {
            gptr_s[0] = 0;
            for (int g = 0; g < G; ++g) gptr_s[g + 1] += gptr_s[g];
            CRWR_CHECK(st, gptr_s[G] == nprod, 11);
        }
        __syncthreads();
        for (int i = tid; i < 32 * G; i += nth) lload_w[i] = lload_s[i];
        for (int i = tid; i <= G; i += nth) gptr_w[i] = gptr_s[i];
        for (int g = warp; g < G; g += W) {
            const int load = (int)lload_s[g * 32 + lane];
            const int rounds = __reduce_max_sync(kFull, load);
            int off = gptr_s[g];
            int ii = 0, c = 0;
            int pp = lane_pos(0, g, G, lane);
            int cntp = pp < no ? (int)cnt_s[pp] : 0;
            int src = pp < no ? bptr_s[pp >> 5] + cptr[pp] : 0;
            T acc = (T)0;
            for (int j = 0; j < rounds; ++j) {
                const bool on = j < load;
                const unsigned m = __ballot_sync(kFull, on);
                if (on) {
                    const int idx = off + __popc(m & lanemask_lt());
                    const unsigned k = tck[src + c];
                    const T w = tcw[src + c];
                    const bool last = c + 1 == cntp;
                    ck_w[idx] = (uint16_t)(k | (last ? kLastFlag : 0u));
                    cw_w[idx] = w;
                    acc = add_rn(acc, mul_rn(q_s[k], w));
                    if (last) {
                        res_s[pp] = acc;
                        acc = (T)0;
                        ++ii; c = 0;
                        pp = lane_pos(ii, g, G, lane);
                        const bool have = chunk_of(ii, g, G) < nb && pp < no;
                        cntp = have ? (int)cnt_s[pp] : 0;
                        src = have ? bptr_s[pp >> 5] + cptr[pp] : 0;
                    } else ++c;
                }
                off += __popc(m);
            }
            __syncwarp();
            int nz = 0;
            for (int i2 = 0, ch; (ch = chunk_of(i2, g, G)) < nb; ++i2) {
                const int q = (ch << 5) + lane;
                const bool act = q < no;
                T v = (T)0;
                if (act) {
                    v = mul_rn(a.scale, res_s[q]);
                    if (q == dpos) v = add_rn(v, a.restart);
                    const T d = sub_rn(q_s[q], v);
                    sq_add(tot.sq, (double)d * (double)d);
                    a.out.vals[slot + q] = v;
                }
                const bool nzv = act && v != (T)0;
                sel.add(nzv, v);
                nz += __popc(__ballot_sync(kFull, nzv));
            }
            if (lane == 0) tot.nnz += (unsigned long long)nz;
        }
        __syncthreads();
        // the table is released; header
        for (int i = tid; i < no; i += nth) { const int sl = oslot[i]; hkey[sl] = -1; hcnt[sl] = 0; }
        if (tid == 0) {
            RowStruct hd;
            hd.off = (long long)poff; hd.slot = slot; hd.no = no; hd.nprod = nprod; hd.nk = nk; hd.dpos = dpos;
            hd.gen = gen; hd.step = step; hd.hot_bytes = (int32_t)L.hot_bytes; hd.groups = G;
            hd.blk_cap = (int32_t)want_blk; hd.slot_cap = (int32_t)want_slot;
            a.rs[row] = hd;
            a.out.row_ptr[row] = slot;
            a.out.row_len[row] = no;
            tot.kept += (unsigned long long)nkt;
            tot.max_len = max(tot.max_len, no);
            tot.max_kept = max(tot.max_kept, nk);
            hot_delta += (long long)L.hot_bytes;
            if (G > 1) ++n_big;
        }
        ++n_built;
        __syncthreads();
    }
    zero_slots(s_slot_base, s_slot_left);
    pdl_launch_dependents();
    if (tid == 0 && n_built) atomicAdd(&st->built_rows, (unsigned)n_built);
    if (tid == 0 && hot_delta) atomicAdd(&st->hot_sum, (unsigned long long)hot_delta);
    if (tid == 0 && n_big) atomicAdd(&st->big_rows, (unsigned)n_big);
    publish_totals(st, tot);
    sel.flush(st);
}

}  // namespace crwr
