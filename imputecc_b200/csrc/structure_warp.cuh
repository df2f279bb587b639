// K2, symbolic part for SHORT rows: the structure of structure.cuh built by ONE WARP per row - used for the FIRST
// build of a walk (every row at once), when the rows average at most ~110 stored entries.
//
// symbolic_kernel gives a row a whole CTA; a short row (a few dozen operands, a few hundred products) leaves most of
// its threads idle behind block-wide barriers: 425 us for the first build of 8,000 such rows.  Here a warp keeps
// everything of its row in a private region of shared memory (~17 KB, sized from the MEAN row: 12-14 rows in flight per
// SM), synchronises with __syncwarp only, and - because the warp walks the row's products in stream order (operand
// column ascending) - the products of an output fall into their summation order by themselves: the place of a product
// is its output's fill level plus its rank among the products of the same output in the same batch of 32 (match_any).
// The products go straight to their final position of the group stream (round tables: how many lanes are still
// active in round j, how many products the rounds before j hold), the hot part of the pool block is assembled in
// shared memory in its pool layout, copied out in 16-byte pieces, and replayed on the spot by the same numeric_stream
// the fast kernel uses.  120 us for the same 8,000 rows.
//
// Rows that exceed the warp's tables (operands, products, outputs, rounds) or would need several groups are handed to
// symbolic_kernel through a second list, untouched.  A SINGLE row is built sooner by a CTA (~15 us against ~35 us for
// one warp: the chain of dependent shared-memory steps is the same, the CTA has more lanes per step), so the short
// lists of later walk steps - the rows whose kept set left their structure - go to symbolic_kernel directly.
#pragma once
#include "structure.cuh"

namespace crwr {

constexpr int kWarpSymMaxWarps = 8;

struct WarpSymShape {
    int kw;             // operand columns (|K+|) a row may have (multiple of 4)
    int pw;             // products
    int now;            // outputs
    int hw;             // hash slots (power of two, >= 1.25 now + 32)
    int rmax;           // rounds of the row's stream
    int warps;          // warps per CTA (0: the kernel is not used)
    uint32_t o_hot, o_q, o_res, o_hkey, o_hcnt, o_hpos, o_olist, o_oslot, o_cnt, o_lbase, o_pslot;
    uint32_t o_kcol, o_kps, o_kpre, o_kval, o_kslot, o_kpos, o_bins, o_rbase, o_rmask;
    uint32_t per_warp;  // bytes of shared memory per warp
    size_t smem;        // dynamic shared memory per CTA (select window included)
};
template <typename T>
__host__ inline WarpSymShape warp_sym_shape(int kw, int pw, int now, int warps) {
    WarpSymShape w;
    const size_t s = sizeof(T);
    w.kw = (kw + 3) & ~3; w.pw = pw; w.now = now; w.warps = warps;
    int hw = 256;
    while (hw < now + now / 4 + 32) hw <<= 1;
    w.hw = hw;
    w.rmax = pw / 32 + w.kw + 8;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += align16(bytes); return (uint32_t)at; };
    w.o_hot = take(struct_layout<T>(now, pw, 1).hot_bytes);
    w.o_q = take((size_t)now * s);
    w.o_res = take((size_t)(now + 32) * s);          // (also the unsorted K+ values; the sums are parked by whole chunks)
    w.o_hkey = take((size_t)hw * 4);
    w.o_hcnt = take((size_t)hw * 4);
    w.o_hpos = take((size_t)hw * 2);
    w.o_olist = take((size_t)now * 2);
    w.o_oslot = take((size_t)now * 2);
    w.o_cnt = take((size_t)now * 2);
    w.o_lbase = take((size_t)now * 2);
    w.o_pslot = take((size_t)(pw > 2 * w.kw + 8 ? pw : 2 * w.kw + 8) * 2);     // (also the unsorted K+ columns)
    w.o_kcol = take((size_t)w.kw * 4);
    w.o_kps = take((size_t)w.kw * 4);
    w.o_kpre = take((size_t)(w.kw + 2) * 4);
    w.o_kval = take((size_t)w.kw * s);
    w.o_kslot = take((size_t)w.kw * 2);
    w.o_kpos = take((size_t)w.kw * 2);
    w.o_bins = take((size_t)(w.kw + 2) * 4);
    w.o_rbase = take((size_t)(w.rmax + 1) * 2);
    w.o_rmask = take((size_t)(w.rmax + 1) * 4);
    w.per_warp = (uint32_t)o;
    w.smem = (size_t)warps * o + align16((size_t)kHistWin * 4);
    return w;
}

template <typename T>
__global__ void __launch_bounds__(kWarpSymMaxWarps * 32) symbolic_warp_kernel(PropArgs<T> a, WarpSymShape ws, int32_t* list2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned int s_wn;
    __shared__ int s_nout[kWarpSymMaxWarps];
    const int tid = threadIdx.x, nth = blockDim.x;
    const int lane = lane_id();
    const int warp = tid >> 5;
    const int W = nth >> 5;
    unsigned char* my = smem_raw + (size_t)warp * ws.per_warp;
    unsigned char* hot = my + ws.o_hot;
    T* q_s = reinterpret_cast<T*>(my + ws.o_q);
    T* res_s = reinterpret_cast<T*>(my + ws.o_res);
    int32_t* hkey = reinterpret_cast<int32_t*>(my + ws.o_hkey);
    int32_t* hcnt = reinterpret_cast<int32_t*>(my + ws.o_hcnt);
    uint16_t* hpos = reinterpret_cast<uint16_t*>(my + ws.o_hpos);
    uint16_t* olist = reinterpret_cast<uint16_t*>(my + ws.o_olist);
    uint16_t* oslot = reinterpret_cast<uint16_t*>(my + ws.o_oslot);
    uint16_t* cnt = reinterpret_cast<uint16_t*>(my + ws.o_cnt);
    uint16_t* lbase = reinterpret_cast<uint16_t*>(my + ws.o_lbase);
    uint16_t* pslot = reinterpret_cast<uint16_t*>(my + ws.o_pslot);
    int32_t* kcol = reinterpret_cast<int32_t*>(my + ws.o_kcol);
    int32_t* kps = reinterpret_cast<int32_t*>(my + ws.o_kps);
    int32_t* kpre = reinterpret_cast<int32_t*>(my + ws.o_kpre);
    T* kval = reinterpret_cast<T*>(my + ws.o_kval);
    uint16_t* kslot = reinterpret_cast<uint16_t*>(my + ws.o_kslot);
    uint16_t* kpos = reinterpret_cast<uint16_t*>(my + ws.o_kpos);
    int32_t* bins = reinterpret_cast<int32_t*>(my + ws.o_bins);
    uint16_t* rbase = reinterpret_cast<uint16_t*>(my + ws.o_rbase);
    uint32_t* rmask = reinterpret_cast<uint32_t*>(my + ws.o_rmask);
    unsigned int* win = reinterpret_cast<unsigned int*>(smem_raw + (size_t)W * ws.per_warp);
    // the unsorted K+ is parked in pslot / res_s until it has been sorted into kcol / kval
    int32_t* kcol_u = reinterpret_cast<int32_t*>(pslot);
    T* kval_u = res_s;
    int* nout = &s_nout[warp];
    const int HM = ws.hw - 1, NOW = ws.now, KW = ws.kw;

    for (int i = tid; i < kHistWin; i += nth) win[i] = 0u;
    for (int i = lane; i < ws.hw; i += 32) { hkey[i] = -1; hcnt[i] = 0; }
    if (tid == 0) s_wn = 0u;
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
    int n_built = 0;
    long long hot_delta = 0;
    if (blockIdx.x == 0 && tid == 0) {
        st->unordered_rows = 1;
        if (!cut_on && n_list > 0) st->order_hazard = 1;      // structures only exist behind a cut (never for positive data)
    }
    const int total_warps = (int)gridDim.x * W;
    // few rows (settled walk): exact reservations; many (first build): a warp reserves for several rows at a time
    const bool chunked = n_list > 2 * total_warps;
    unsigned long long slot_base = 0ull, slot_left = 0ull, pool_base = 0ull, pool_left = 0ull;    // (uniform over the warp)
    const int32_t* __restrict__ pcols = a.p_cols;
    const T* __restrict__ pvals = a.p_vals;

    auto insert = [&](int32_t j) -> int {
        int sl = (int)sym_hash(j, HM);
        for (;;) {
            const int32_t k = hkey[sl];
            if (k == j) return sl;
            if (k == -1) {
                const int32_t o = atomicCAS(&hkey[sl], -1, j);
                if (o == -1) {
                    const int idx = atomicAdd(nout, 1);
                    if (idx < NOW) olist[idx] = (uint16_t)sl;
                    return sl;
                }
                if (o == j) return sl;
            }
            sl = (sl + 1) & HM;
        }
    };
    auto zero_slots = [&](unsigned long long base, unsigned long long left) {
        for (unsigned long long i = lane; i < left && base + i < (unsigned long long)a.slot_cap; i += 32) {
            a.in.vals[a.fixed_base + (long long)(base + i)] = (T)0;
            a.out.vals[a.fixed_base + (long long)(base + i)] = (T)0;
        }
    };
    auto release_table = [&](int no) {                          // (warp-wide)
        if (no > NOW) { for (int i = lane; i < ws.hw; i += 32) { hkey[i] = -1; hcnt[i] = 0; } }
        else for (int i = lane; i < no; i += 32) { const int sl = olist[i]; hkey[sl] = -1; hcnt[sl] = 0; }
        __syncwarp();
    };
    auto hand_over = [&](int row) { if (lane == 0) list2[atomicAdd(&st->slow2_count, 1u)] = row; };

    if (cut_on)
    for (int li = warp * (int)gridDim.x + (int)blockIdx.x; li < n_list; li += total_warps) {
        const int row = a.row_list_count ? __ldg(a.row_order + li) : li;
        const int32_t dcol = a.row_begin + row;
        const RowStruct old = a.rs[row];
        const long long ptr = a.in.row_ptr[row];
        const int len = a.in.row_len[row];
        // ---- P1: K+ (unordered), then sorted by column (rank by counting)
        int nk = 0, nkt = 0;
        for (int b = 0; b < len; b += 128) {
            T v[4];
            int32_t c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = b + u * 32 + lane;
                v[u] = (T)0; c[u] = 0;
                if (e < len) { v[u] = __ldg(a.in.vals + ptr + e); c[u] = __ldg(a.in.cols + ptr + e); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool keep = v[u] > cut_lo;
                const bool kt = v[u] > cut;
                const unsigned m = __ballot_sync(kFull, keep);
                const int pos = nk + __popc(m & lanemask_lt());
                if (keep && pos < KW) { kcol_u[pos] = c[u]; kval_u[pos] = kt ? v[u] : (T)0; }
                nk += __popc(m);
                nkt += __popc(__ballot_sync(kFull, kt));
            }
        }
        if (nk > KW) { hand_over(row); __syncwarp(); continue; }
        for (int i = nk + lane; i < ((nk + 3) & ~3); i += 32) kcol_u[i] = 0x7fffffff;
        __syncwarp();
        for (int i = lane; i < nk; i += 32) {
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
        __syncwarp();
        int nprod = 0;
        for (int b = 0; b < nk; b += 32) {
            const int i = b + lane;
            const int l = i < nk ? kpre[i + 1] : 0;
            int incl = l;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            if (i < nk) kpre[i + 1] = nprod + incl;
            nprod += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) { kpre[0] = 0; *nout = 0; }
        __syncwarp();
        if (nprod > ws.pw) { hand_over(row); continue; }
        // ---- P2: the output set and the products per output (a lane per product, 128 products per batch)
        for (int i = lane; i < nk; i += 32) kslot[i] = (uint16_t)insert(kcol[i]);
        int dslot = 0;
        if (lane == 0) dslot = insert(dcol);
        dslot = __shfl_sync(kFull, dslot, 0);
        __syncwarp();
        {
            int k = 0;
            for (int i0 = 0; i0 < nprod; i0 += 128) {
                int32_t j[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    j[u] = -1;
                    if (i < nprod) {
                        while (kpre[k + 1] <= i) ++k;
                        j[u] = __ldg(pcols + kps[k] + (i - kpre[k]));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    if (i < nprod) {
                        int sl = 0;
                        if (*(volatile int*)nout <= NOW) {
                            CRWR_CHECK(st, j[u] >= 0, 5);
                            sl = insert(j[u]);
                            atomicAdd(&hcnt[sl], 1);
                        }
                        pslot[i] = (uint16_t)sl;
                    }
                }
            }
        }
        __syncwarp();
        const int no = *(volatile int*)nout;
        if (no > NOW) { release_table(no); hand_over(row); continue; }
        // ---- P3: counting sort by count, descending
        for (int i = lane; i <= nk; i += 32) bins[i] = 0;
        __syncwarp();
        for (int o = lane; o < no; o += 32) atomicAdd(&bins[hcnt[olist[o]]], 1);
        __syncwarp();
        {
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
        __syncwarp();
        for (int o = lane; o < no; o += 32) {
            const int sl = olist[o];
            const int c = hcnt[sl];
            const int pos = atomicAdd(&bins[c], 1);
            oslot[pos] = (uint16_t)sl;
            cnt[pos] = (uint16_t)c;
            hpos[sl] = (uint16_t)pos;
            hcnt[sl] = 0;                   // from here on: the output's fill level
        }
        __syncwarp();
        // ---- the lanes' shares (chunk by chunk, alternating direction) and the round tables of the stream
        const int nb = (no + 31) >> 5;
        int load = 0;
        for (int ii = 0; ii < nb; ++ii) {
            const int pp = lane_pos(ii, 0, 1, lane);
            if (pp < no) { lbase[pp] = (uint16_t)load; load += (int)cnt[pp]; }
        }
        const int rounds = __reduce_max_sync(kFull, load);
        const StructLayout L = struct_layout<T>(no, nprod, 1);
        if (rounds > ws.rmax || no > a.small_no || L.hot_bytes > (uint32_t)a.small_hot) {
            release_table(no); hand_over(row); continue;        // (a row of several groups is the CTA kernel's)
        }
        {
            int run = 0;
            for (int j = 0; j < rounds; ++j) {
                const unsigned m = __ballot_sync(kFull, load > j);
                if (lane == 0) { rmask[j] = m; rbase[j] = (uint16_t)run; }
                run += __popc(m);
            }
            CRWR_CHECK(st, run == nprod, 8);
        }
        // ---- reservations: fixed slot in both iterate buffers, pool block (lane 0; chunked while many rows are built)
        const unsigned long long no4 = (unsigned long long)((no + 3) & ~3);     // slots start on 16-byte boundaries (bulk copies)
        if (no4 > slot_left) {
            zero_slots(slot_base, slot_left);
            const unsigned long long sz = (chunked && (unsigned long long)a.slot_chunk > no4) ? (unsigned long long)a.slot_chunk : no4;
            unsigned long long b = 0ull;
            if (lane == 0) b = atomicAdd(&st->slot_cursor, sz);
            slot_base = __shfl_sync(kFull, b, 0);
            slot_left = sz;
        }
        const unsigned long long slot_rel = slot_base;
        slot_base += no4; slot_left -= no4;
        if ((unsigned long long)L.bytes > pool_left) {
            const unsigned long long sz = (chunked && (unsigned long long)a.pool_chunk > (unsigned long long)L.bytes) ? (unsigned long long)a.pool_chunk : (unsigned long long)L.bytes;
            unsigned long long b = 0ull;
            if (lane == 0) b = atomicAdd(&st->pool_cursor, sz);
            pool_base = __shfl_sync(kFull, b, 0);
            pool_left = sz;
        }
        const unsigned long long poff = pool_base;
        pool_base += (unsigned long long)L.bytes; pool_left -= (unsigned long long)L.bytes;
        // the row's old fixed slot (if any) is abandoned: no stale value may be left for a whole-array select pass.  The copy
        // in this step's output buffer now; the one in the input buffer (the row's operand) unless the row is deferred to the
        // large-row kernel, which still reads it (and clears it afterwards).
        const bool had_slot = old.gen == gen && old.no > 0;
        if (had_slot) {
            for (int i = lane; i < old.no; i += 32) a.out.vals[old.slot + i] = (T)0;
            if (lane == 0) { a.rs[row].gen = 0u; hot_delta -= (long long)old.hot_bytes; }
        }
        if (slot_rel + no4 > (unsigned long long)a.slot_cap || poff + L.bytes > a.pool_cap) {
            // slot region or pool exhausted: the row goes without a structure (large-row kernel), this step and - as the
            // cursors only grow - every later one; struct_full tells the host a larger workspace would be faster
            if (lane == 0) {
                st->struct_full = 1;
                a.fb_rows[atomicAdd(&st->fb_count, 1u)] = row;
                tot.max_kept = max(tot.max_kept, nk);
            }
            zero_slots(slot_rel, no4);          // the slot entries it did get stay empty
            release_table(no);
            continue;
        }
        const long long slot = a.fixed_base + (long long)slot_rel;
        if (lane < (int)no4 - no) { a.in.vals[slot + no + lane] = (T)0; a.out.vals[slot + no + lane] = (T)0; }
        if (had_slot) for (int i = lane; i < old.no; i += 32) a.in.vals[old.slot + i] = (T)0;
        // ---- the hot part of the pool block, assembled in shared memory: shares, K+ bits, previous values in structure order
        uint16_t* lload_s = reinterpret_cast<uint16_t*>(hot + L.lload);
        int32_t* gptr_s = reinterpret_cast<int32_t*>(hot + L.gptr);
        uint32_t* kbits_s = reinterpret_cast<uint32_t*>(hot + L.kbits);
        uint16_t* ck_s = reinterpret_cast<uint16_t*>(hot + L.ck);
        T* cw_s = reinterpret_cast<T*>(hot + L.cw);
        lload_s[lane] = (uint16_t)load;
        if (lane == 0) { gptr_s[0] = 0; gptr_s[1] = nprod; }
        for (int i = lane; i < nb; i += 32) kbits_s[i] = 0u;
        for (int i = lane; i < no; i += 32) q_s[i] = (T)0;
        __syncwarp();
        for (int i = lane; i < nk; i += 32) {
            const int pos = hpos[kslot[i]];
            kpos[i] = (uint16_t)pos;
            q_s[pos] = kval[i];
            atomicOr(&kbits_s[pos >> 5], 1u << (pos & 31));
        }
        const int dpos = hpos[dslot];
        __syncwarp();
        // ---- P4: the products to their places in the stream.  Batches of 32 in stream order: a product's rank in its
        // output's list = the output's fill level + the number of lower lanes of the batch with the same output
        {
            int k = 0;
            T w_next = (T)0;
            if (lane < nprod) {
                while (kpre[k + 1] <= lane) ++k;
                w_next = __ldg(pvals + kps[k] + (lane - kpre[k]));
            }
            int k_next = k;
            for (int i0 = 0; i0 < nprod; i0 += 32) {
                const int i = i0 + lane;
                const bool on = i < nprod;
                const T w = w_next;
                k = k_next;
                const int in = i + 32;
                if (in < nprod) {
                    while (kpre[k_next + 1] <= in) ++k_next;
                    w_next = __ldg(pvals + kps[k_next] + (in - kpre[k_next]));
                }
                const unsigned sl = on ? (unsigned)pslot[i] : (0x10000u | (unsigned)lane);
                const unsigned peers = __match_any_sync(kFull, sl);
                const int rank = __popc(peers & lanemask_lt());
                int c = 0, pos = 0;
                if (on) { pos = (int)hpos[sl]; c = hcnt[sl] + rank; }
                __syncwarp();
                if (on && rank == 0) hcnt[sl] += __popc(peers);
                __syncwarp();
                if (on) {
                    const int seq = (int)lbase[pos] + c;
                    const int owner = ((pos >> 5) & 1) ? 31 - (pos & 31) : (pos & 31);
                    const int at = (int)rbase[seq] + __popc(rmask[seq] & ((1u << owner) - 1u));
                    CRWR_CHECK(st, at >= 0 && at < nprod && seq < rounds, 9);
                    ck_s[at] = (uint16_t)((unsigned)kpos[k] | (c + 1 == (int)cnt[pos] ? kLastFlag : 0u));
                    cw_s[at] = w;
                }
            }
        }
        __syncwarp();
        // ---- P5: pool block (hot part in 16-byte pieces, output columns), the slot's columns in both buffers
        unsigned char* blk = a.pool + poff;
        {
            const uint4* src = reinterpret_cast<const uint4*>(hot);
            uint4* dst = reinterpret_cast<uint4*>(blk);
            for (int i = lane; i < (int)(L.hot_bytes >> 4); i += 32) dst[i] = src[i];
        }
        int32_t* ocols_w = reinterpret_cast<int32_t*>(blk + L.ocols);
        for (int pp = lane; pp < no; pp += 32) {
            const int32_t col = hkey[oslot[pp]];
            ocols_w[pp] = col;
            a.in.cols[slot + pp] = col;
            a.out.cols[slot + pp] = col;
        }
        // ---- P6: the numeric pass, exactly as the fast kernel will replay it
        numeric_stream<T, kStaged>(a, hot, no, nprod, 1, 0, dpos, q_s, res_s, slot, sel, tot);
        __syncwarp();
        for (int i = lane; i < no; i += 32) { const int sl = oslot[i]; hkey[sl] = -1; hcnt[sl] = 0; }
        if (lane == 0) {
            RowStruct hd;
            hd.off = (long long)poff; hd.slot = slot; hd.no = no; hd.nprod = nprod; hd.nk = nk; hd.dpos = dpos;
            hd.gen = gen; hd.step = step; hd.hot_bytes = (int32_t)L.hot_bytes; hd.groups = 1;
            hd.blk_cap = (int32_t)L.bytes; hd.slot_cap = (int32_t)no4;
            a.rs[row] = hd;
            a.out.row_ptr[row] = slot;
            a.out.row_len[row] = no;
            tot.kept += (unsigned long long)nkt;
            tot.max_len = max(tot.max_len, no);
            tot.max_kept = max(tot.max_kept, nk);
            hot_delta += (long long)L.hot_bytes;
        }
        ++n_built;
        __syncwarp();
    }
    zero_slots(slot_base, slot_left);
    pdl_launch_dependents();
    if (lane == 0 && n_built) atomicAdd(&st->built_rows, (unsigned)n_built);
    if (lane == 0 && hot_delta) atomicAdd(&st->hot_sum, (unsigned long long)hot_delta);
    publish_totals(st, tot);
    sel.flush(st);
}

}  // namespace crwr
