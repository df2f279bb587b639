// C ABI of libcrwr_b200.so (see include/crwr_b200.h).  Host-side glue only: workspace layout,
// kernel launches, status copies.  No device allocation, no CPU compute path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <cmath>
#include <utility>
#include <map>

#include "common.cuh"
#include "select.cuh"
#include "propagate.cuh"
#include "propagate_group.cuh"
#include "tail.cuh"
#include "structure_warp.cuh"
#include "transition.cuh"
#include "finish.cuh"
#include "peer.cuh"
#include "binweights.cuh"

using namespace crwr;

static thread_local std::string g_err;
static long long g_launches = 0;

#define CRWR_FAIL(code, ...)                               \
    do {                                                   \
        char _b[512];                                      \
        snprintf(_b, sizeof(_b), __VA_ARGS__);             \
        g_err = _b;                                        \
        return (code);                                     \
    } while (0)

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) CRWR_FAIL(CRWR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define LAUNCH_CHECK()                                                                          \
    do {                                                                                        \
        ++g_launches;                                                                           \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) CRWR_FAIL(CRWR_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
    } while (0)

namespace {

constexpr int kDenseWarpsPerCta = 4;
constexpr int kOrderStep = 4;       // first walk step that hands rows out longest first (row lengths have settled by then)

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
double env_double(const char* name, double dflt) {
    const char* e = getenv(name);
    return e ? atof(e) : dflt;
}

// Launch with programmatic stream serialization (see pdl_wait in common.cuh) when `pdl`, else the ordinary way.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// The dynamic shared-memory limit is an attribute of the kernel, not of a handle: walkers with different shapes share
// it, so it only ever grows (every launch states its own size).
std::map<const void*, size_t> g_smem_attr;
std::mutex g_smem_attr_mutex;
template <typename K>
cudaError_t raise_smem_attr(K kernel, size_t smem) {
    std::lock_guard<std::mutex> lock(g_smem_attr_mutex);
    size_t& cur = g_smem_attr[(const void*)kernel];
    if (smem <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) cur = smem;
    return e;
}

struct Layout {
    size_t state, trace, p_indptr, p_cols, p_vals;
    size_t it_cols[2], it_vals[2], it_ptr[2], it_len[2];
    size_t fb_rows, row_order, scan_tmp, xbuf, cand, merged, gathered, counts, counts_scan, total;
    size_t d_acc, d_oldv, d_ord, d_tbits, d_kbits;
    size_t k1_row_cnt, k1_col_cnt, k1_col_ptr, k1_col_fill, k1_csc_row, k1_csc_val, k1_scale, k1_iso, k1_row_val;
    size_t k1_bad;
    size_t rs, slow_list, slow_list2, pool, sym_tmp;
    size_t pool_bytes;
    long long fixed_base, slot_cap; // first entry / entries of the fixed-slot region behind the per-step entries of both iterate buffers
    int sym_tmp_cap;                // products per staging region of the symbolic kernel (with the most CTAs it is launched with)
    size_t sym_tmp_total;           // bytes of all staging regions
    long long sym_tmp_max;          // a row has at most this many products (nnz of the transition matrix)
    size_t bytes;
    long long cand_cap;
    long long dense_warps;
    long long words;
};

size_t bump(size_t& off, size_t bytes) {
    size_t at = (off + 255) & ~(size_t)255;
    off = at + bytes;
    return at;
}

size_t elem(int dtype) { return dtype == CRWR_F64 ? 8 : 4; }

bool make_layout(const crwr_config& c, Layout& L) {
    const size_t s = elem(c.dtype);
    const long long n = c.n_contigs;
    const long long rows = c.row_end - c.row_begin;
    const long long pcap = c.transition_nnz_cap;
    const long long cap = c.iterate_cap;
    size_t off = 0;
    L.state = bump(off, sizeof(WalkState));
    L.trace = bump(off, sizeof(crwr_step_record) * kMaxTrace);
    L.p_indptr = bump(off, 4 * (size_t)(n + 1));
    L.p_cols = bump(off, 4 * (size_t)pcap);
    L.p_vals = bump(off, s * (size_t)pcap);
    // iterate buffers: [cap per-step entries | slot_cap fixed-slot entries (rows with a cached structure, structure.cuh)]
    // The fixed region starts on a 16-entry boundary (bulk copies) and has as many entries as the per-step region.
    L.fixed_base = (cap + 15) & ~15ll;
    L.slot_cap = L.fixed_base;
    for (int b = 0; b < 2; ++b) {
        L.it_cols[b] = bump(off, 4 * (size_t)(L.fixed_base + L.slot_cap));
        L.it_vals[b] = bump(off, s * (size_t)(L.fixed_base + L.slot_cap));
        L.it_ptr[b] = bump(off, 8 * (size_t)rows);
        L.it_len[b] = bump(off, 4 * (size_t)rows);
    }
    L.fb_rows = bump(off, 4 * (size_t)rows);
    L.row_order = bump(off, 4 * (size_t)rows);
    L.scan_tmp = bump(off, 8 * (size_t)kScanMaxTiles);
    // exchange buffer: [extras (kXExtras) | level-0 histogram | level-1 histogram], all uint64
    L.xbuf = bump(off, 8 * (size_t)(kXExtras + kSelectLevels * kSelectBins));
    // candidate block [count, successor, keys...]; sharded walks use a small fixed block (all-gathered)
    if (c.n_shards > 1) L.cand_cap = CRWR_SHARD_CAND;
    else L.cand_cap = cap / 16 > 65536 ? cap / 16 : 65536;
    L.cand = bump(off, 8 * (size_t)(kCandHeader + L.cand_cap));
    L.merged = bump(off, 8 * (size_t)(c.n_shards > 1 ? (long long)c.n_shards * L.cand_cap : 1));
    L.gathered = bump(off, 8 * (size_t)(c.n_shards > 1 ? (long long)c.n_shards * kPeerCandWords : 1));
    L.counts = bump(off, 4 * (size_t)(rows + 1));
    L.counts_scan = bump(off, 8 * (size_t)(rows + 1));
    L.total = bump(off, 8);
    // large-row kernel scratch: as many warps as fit in ~1 GiB, between 8 and 592
    L.words = (n + 31) / 32;
    const size_t per_warp = (size_t)n * (2 * s + 4) + (size_t)L.words * 8;
    long long dw = (long long)(((size_t)1 << 30) / (per_warp ? per_warp : 1));
    if (dw > 592) dw = 592;
    if (dw < 8) dw = 8;
    dw = (dw / kDenseWarpsPerCta) * kDenseWarpsPerCta;
    L.dense_warps = dw;
    L.d_acc = bump(off, s * (size_t)n * dw);
    L.d_oldv = bump(off, s * (size_t)n * dw);
    L.d_ord = bump(off, 4 * (size_t)n * dw);
    L.d_tbits = bump(off, 4 * (size_t)L.words * dw);
    L.d_kbits = bump(off, 4 * (size_t)L.words * dw);
    // transition-build scratch
    L.k1_row_cnt = bump(off, 4 * (size_t)(n + 1));
    L.k1_col_cnt = bump(off, 4 * (size_t)(n + 1));
    L.k1_col_ptr = bump(off, 4 * (size_t)(n + 1));
    L.k1_col_fill = bump(off, 4 * (size_t)n);
    L.k1_csc_row = bump(off, 4 * (size_t)pcap);
    L.k1_csc_val = bump(off, 8 * (size_t)pcap);
    L.k1_scale = bump(off, 8 * (size_t)n);
    L.k1_iso = bump(off, 4 * (size_t)n);
    L.k1_row_val = bump(off, 8 * (size_t)pcap);
    L.k1_bad = bump(off, 8);
    // cached row structures (structure.cuh): headers, the fast kernel's hand-over list, the pool, the symbolic kernel's
    // staging regions.  A row costs about (2 + s) bytes per product plus ~10 per output; rebuilt blocks are never
    // reclaimed during a walk, so the pool is sized well above the working set of the largest walk the iterate buffers
    // can hold (running out is a capacity error: the host regrows the workspace).
    L.rs = bump(off, sizeof(RowStruct) * (size_t)rows);
    L.slow_list = bump(off, 4 * (size_t)rows);
    L.slow_list2 = bump(off, 4 * (size_t)rows);
    {
        long long pb = (long long)(c.dtype == CRWR_F64 ? 80 : 48) * cap;
        if (pb < (64ll << 20)) pb = 64ll << 20;
        if (const char* e = getenv("CRWR_POOL_BYTES")) { const long long v = atoll(e); if (v >= 4096) pb = v; }
        L.pool_bytes = (size_t)pb;
    }
    L.pool = bump(off, L.pool_bytes);
    {
        long long tc = pcap < 16384 ? pcap : 16384;         // a row's products are at most nnz(P)
        if (tc < 1024) tc = 1024;
        tc = (tc + 7) & ~7ll;
        L.sym_tmp_cap = (int)tc;
        const long long ctas = rows < kSymMaxCtas ? rows : kSymMaxCtas;
        L.sym_tmp_total = (size_t)ctas * sym_tmp_bytes((int)tc, s);
        L.sym_tmp_max = pcap;
        L.sym_tmp = bump(off, L.sym_tmp_total);
    }

    L.bytes = (off + 255) & ~(size_t)255;
    return true;
}

}  // namespace

struct crwr_handle_s {
    crwr_config cfg;
    Layout L;
    unsigned char* ws;
    int sm_count;
    long long p_nnz;
    bool have_p;
    bool walking;
    int steps_enqueued;
    PropShape shape;                // run-time sizing of the warp-per-row propagate kernel (walk step 0)
    int grid;                       // its grid size
    GroupShape gshape;              // run-time sizing of the row-group propagate kernel (default)
    int ggrid = 0;                  // its persistent grid size
    int prop_kind = 1;              // 1: row-group kernel (default), 0: warp-per-row kernel (CRWR_PROP=warp)
    bool order_valid = false;       // row_order holds the longest-first order of this walk (built before step kOrderStep)
    bool use_order = true;          // CRWR_ROW_ORDER=0: rows in index order
    int pdl_forced = -1;            // CRWR_PDL=0/1: programmatic dependent launch forced off/on (-1: by row length, use_pdl)
    int max_smem_optin;
    double rp, perc;
    WalkState* h_state = nullptr;               // pinned host mirror of the device state (polls)
    long long hint_len = -1, hint_kept = -1;    // last sizing hints (configure_shape is skipped when unchanged)
    long long hint_len_eff = 128;               // expected longest product row
    PeerInfo peer = {nullptr, 0, 0};           // NVLink peer exchange (sharded walk), see peer.cuh
    bool peer_attached = false;
    unsigned long long* peer_block = nullptr;   // this rank's symmetric block
    unsigned long long peer_epoch_base = 0;
    bool hist0_fused = false;                   // sharded step: level-0 histogram already filled by propagate
    int struct_mode = 1;                        // CRWR_STRUCT: 0 no cached row structures (row-group kernel every step), 1 when
                                                // the rows are long enough to pay for them (default), 2 always
    long long struct_min_row = 176;             // ... mean stored entries per row (at the last poll before the build: walk step 1)
                                                // from which they pay (CRWR_STRUCT_MIN_ROW; measured crossover 161 / 191, profiles/bench_r2.md)
    long long mean_row = 0;                     // mean stored entries per row at the last poll
    bool structs_active = false;                // this walk's decision, taken when the first structure step is queued
    bool redo_hist = false;                     // sharded walk: the next step repeats one whose window missed (histogram select)
    int win_step = 6;                           // first walk step whose cut comes from the window select (CRWR_WIN_STEP; 0: never)
    bool scan_push = true;                      // sharded histogram steps exchange level 0 (+1) by pushing (CRWR_SCAN_PUSH=0: pull)
    bool sharded_window_ok = false;             // sharded walk: the last poll saw a window narrow enough for the push regions
    int win_step_sharded = 9;
    int win_hold_until = 0;                     // host policy (crwr_hold_window_select): no window-select step before this walk step
    float alpha = 0.5f;                         // K+ = entries above alpha x cut (CRWR_STRUCT_ALPHA)
    float alpha_build = 0.4f;                   // the same for the first build (CRWR_STRUCT_ALPHA_BUILD, default 0.8 x alpha): the rows are
                                                // still growing then, a wider K+ keeps a quarter of them from being rebuilt in the next steps
    float win_scale = 2.5f, win_min = 0.004f, win_max = 0.15f, win_rel_max = 0.10f;
    bool dense_always = false;                  // a step deferred rows while no large-row kernel was queued: always queue it
    bool fb_seen = false;                       // this walk has deferred rows before (polled)
    SymShape sym = {0, 0, 0, 0, 0, 0, 0, 0};    // sizing of the symbolic kernel (structure.cuh)
    int sym_grid = 0;
    WarpSymShape wsym = {};                     // sizing of the warp-per-row symbolic kernel (structure_warp.cuh); warps 0: not used
    int wsym_grid = 0;
    int warp_sym = 1;                           // CRWR_WARP_SYM=0: short rows are built by the CTA kernel too
    int fast_no = 0, fast_grid = 0;             // fast kernel: outputs per row it holds (0: kernel unusable), persistent grid
    FastShape fshape = {0, 0, 0, 0};
    size_t big_smem = 0;                        // big-row kernel (structure.cuh): shared memory, persistent grid
    int big_grid = 0;
    bool big_seen = false;                      // this walk has long rows (polled): the big-row kernel is queued
    long long avg_hot = 0;                      // mean pool bytes per row of this walk's structures (polled; 0: unknown)
    bool fast_retune = false;                   // avg_hot moved: the next crwr_set_row_hints sizes the fast kernel again
    int build_step = 2;                         // first walk step that runs the symbolic kernel (CRWR_BUILD_STEP, >= 2)
    bool dense_launched = true;                 // the step being enqueued has a large-row kernel behind its propagation
    bool bitmaps_zeroed = false;                // large-row kernel bitmaps cleared (first crwr_walk_init, caller's stream)
    bool profile = false;                       // CUDA events around every main propagate launch
    std::vector<cudaEvent_t> ev;                // start/end pairs
    std::vector<cudaEvent_t> marks;             // profile mode: step start + one event after each kernel of the step
    std::vector<int> mark_ids;
    double phase_ms[4] = {0, 0, 0, 0};
    std::vector<double> pair_ms;                // durations of the propagation launches of the last profiled walk, step by step
    bool timeline = false;                      // CRWR_TIMELINE=1 (developer aid): an event after every kernel of a profiled walk,
    std::vector<std::pair<cudaEvent_t, const char*>> tl;   // durations printed to stderr by crwr_profile_read
    template <typename U> U* at(size_t off) const { return reinterpret_cast<U*>(ws + off); }
};

namespace {

// Pinned host mirrors of the walk state are pooled: cudaMallocHost / cudaFreeHost cost about a
// millisecond each, more than a whole small walk.
std::vector<WalkState*> g_pinned_pool;
std::mutex g_pinned_mutex;
WalkState* acquire_pinned_state() {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    if (!g_pinned_pool.empty()) {
        WalkState* p = g_pinned_pool.back();
        g_pinned_pool.pop_back();
        return p;
    }
    WalkState* p = nullptr;
    if (cudaMallocHost((void**)&p, sizeof(WalkState)) != cudaSuccess) return nullptr;
    return p;
}
void release_pinned_state(WalkState* p) {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    g_pinned_pool.push_back(p);
}

int round_up(long long x, int q) { return (int)(((x + q - 1) / q) * q); }

// Sizes the accumulator table / kept list / staging buffers from the caller's row-length hints.
template <typename T>
int configure_shape(crwr_handle_s* h, long long max_row_len, long long max_kept) {
    const long long n = h->cfg.n_contigs;
    long long limit = round_up(max_row_len + max_row_len / 4 + 48, 32);
    if (limit < 256) limit = 256;
    if (limit > n + 32) limit = round_up(n + 32, 32);
    long long kmax = round_up(max_kept + max_kept / 4 + 32, 32);
    if (kmax < 64) kmax = 64;
    if (kmax > n) kmax = round_up(n, 32);
    const double avg_deg = h->p_nnz > 0 ? (double)h->p_nnz / (double)n : 8.0;
    // staging buffer: about 16 P rows per batch - enough rounds to cover the cp.async latency of the next one
    long long scap = round_up((long long)(16.0 * avg_deg), 64);
    if (scap < 128) scap = 128;
    if (scap > 768) scap = 768;
    PropShape sh;
    for (;;) {
        long long H = round_up(limit + limit / 3 + 32, 32);
        if (H > 65535) { H = 65504; limit = (H * 3) / 4; }
        sh.H = (int)H; sh.limit = (int)limit; sh.kmax = (int)kmax; sh.scap = (int)scap;
        sh.per_warp = prop_smem_per_warp<T>(sh.H, sh.limit, sh.kmax, sh.scap);
        if ((long long)sh.per_warp <= (long long)h->max_smem_optin) break;
        // too big for one SM: shrink (oversized rows are deferred to the large-row kernel)
        if (kmax > 1024 && kmax * 2 > limit) kmax = round_up(kmax / 2, 32);
        else limit = round_up(limit * 3 / 4, 32);
    }
    int w = (int)((57344 - kHistWin * sizeof(unsigned int)) / sh.per_warp);
    if (w < 1) w = 1;
    if (w > 8) w = 8;
    while (w & (w - 1)) --w;        // power of two: block-wide scans are templated on the CTA size
    sh.warps = w;
    const size_t smem = sh.per_warp * (size_t)w + kHistWin * sizeof(unsigned int);
    cudaError_t e = raise_smem_attr(propagate_kernel<T>, smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, propagate_kernel<T>, w * 32, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    h->shape = sh;
    h->grid = per_sm * h->sm_count;
    return 0;
}

template <typename T, int G>
int group_kernel_attr(size_t smem, int threads, int* per_sm, int) {
    cudaError_t e = raise_smem_attr(propagate_group_kernel<T, G>, smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, propagate_group_kernel<T, G>, threads, smem);
    return (int)e;
}

// Sizes the symbolic and the fast kernel (structure.cuh) from the caller's hints: outputs and operand columns per row.
template <typename T>
int configure_struct_shape(crwr_handle_s* h, long long max_row_len, long long max_kept) {
    const long long n = h->cfg.n_contigs;
    const long long rows = h->cfg.row_end - h->cfg.row_begin;
    const size_t win_bytes = kHistWin * sizeof(unsigned int);
    // (a structure covers what the superset K+ reaches, more than the stored row the hint was taken from)
    long long no = round_up(max_row_len + max_row_len * 6 / 10 + 32, 32);
    if (no < 128) no = 128;
    if (no > n + 1) no = round_up(n + 1, 32);
    // (K+ holds the entries above alpha x cut: up to about twice the kept ones the hint counts)
    long long km = round_up(max_kept * 22 / 10 + 16, 32);
    if (km < 64) km = 64;
    if (km > no) km = no;
    SymShape sh;
    // shared-memory staging for rows of up to ~16 products per output (longer rows stage in global memory)
    const double avg_deg = h->p_nnz > 0 ? (double)h->p_nnz / (double)n : 8.0;
    for (;;) {
        if (no > 32736) no = 32736;            // an output position travels in 15 bits of a stream entry (bit 15 flags a list end)
        if (km > no) km = no;
        long long H = 512;
        while (H < 2 * no + 64) H <<= 1;
        long long ts = round_up((long long)((double)km * avg_deg * 1.25) + 256, 64);
        if (ts > 8192) ts = 8192;
        if (ts > h->L.sym_tmp_cap) ts = h->L.sym_tmp_cap;
        sh.H = (int)H; sh.no_max = (int)no; sh.km = (int)km; sh.tmp_smem = (int)ts;
        sh.smem = sym_smem_bytes(sh.H, sh.no_max, sh.km, sh.tmp_smem, sizeof(T));
        // too big for one SM: first fewer products staged in shared memory (the rest stage in global memory), only then
        // fewer outputs - rows beyond them go to the large-row kernel, 100 us per step
        while (sh.smem + 1024 > (size_t)h->max_smem_optin && sh.tmp_smem > 2048) {
            sh.tmp_smem = (int)round_up(sh.tmp_smem * 3 / 4, 64);
            sh.smem = sym_smem_bytes(sh.H, sh.no_max, sh.km, sh.tmp_smem, sizeof(T));
        }
        if (sh.smem + 1024 <= (size_t)h->max_smem_optin) break;
        no = round_up(no * 3 / 4, 32);
    }
    sh.tmp_cap = h->L.sym_tmp_cap;
    sh.threads = env_int("CRWR_SYM_THREADS", 256);
    sh.threads_build = env_int("CRWR_SYM_THREADS_BUILD", 128);
    if (sh.threads < 32 || sh.threads > 256 || (sh.threads & 31)) sh.threads = 256;
    if (sh.threads_build < 32 || sh.threads_build > 256 || (sh.threads_build & 31)) sh.threads_build = 128;
    cudaError_t e = raise_smem_attr(symbolic_kernel<T>, sh.smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, symbolic_kernel<T>, sh.threads_build, sh.smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm == 1) { sh.threads = 512; sh.threads_build = 512; }      // long rows: shared memory admits one CTA per SM, so it gets more threads
    long long grid = (long long)per_sm * h->sm_count;
    const long long max_ctas = rows < kSymMaxCtas ? rows : kSymMaxCtas;
    if (grid > max_ctas) grid = max_ctas;
    if (grid < 1) grid = 1;
    // the staging regions are shared out among the CTAs actually launched: long rows run one CTA per SM and may have
    // several times the products a region holds when every SM carries eight CTAs
    {
        long long cap = sh.tmp_cap;
        while (cap * 2 <= h->L.sym_tmp_max * 2 && (size_t)grid * sym_tmp_bytes((int)(cap * 2), sizeof(T)) <= h->L.sym_tmp_total &&
               cap * 2 <= (1ll << 22))
            cap *= 2;
        if (cap > h->L.sym_tmp_max) cap = (h->L.sym_tmp_max + 7) & ~7ll;
        if (cap > sh.tmp_cap) sh.tmp_cap = (int)cap;
    }
    h->sym = sh;
    h->sym_grid = (int)grid;
    // short rows, first build: a warp per row (structure_warp.cuh).  The tables are sized from the MEAN row as the last poll
    // saw it (walk step 1; rows grow by a third until the build at step 3): 2.2 x its stored entries, a quarter of that in
    // operands, products for the mean length of a P row - a warp's shared memory decides how many rows are in flight; whatever
    // is larger goes to the CTA kernel above (at a mean above ~110 most rows would, after the warp has done half the work).  The CTA size that keeps most warps on an SM.
    h->wsym.warps = 0;
    if (h->warp_sym && h->mean_row > 0 && h->mean_row <= 110) {
        long long now = round_up((long long)(2.2 * (double)h->mean_row) + 32, 32);
        if (now < 128) now = 128;
        long long kw = round_up(now / 4, 32);
        if (kw < 64) kw = 64;
        long long pw = round_up((long long)((double)kw * avg_deg * 1.1) + 64, 64);
        if (pw < 512) pw = 512;
        if (pw > 2048) pw = 2048;
        int best_w = 0, best_warps = 0, best_per_sm = 0;
        for (int w = 2; w <= kWarpSymMaxWarps; ++w) {
            const WarpSymShape c = warp_sym_shape<T>((int)kw, (int)pw, (int)now, w);
            if (c.smem + 1024 > (size_t)h->max_smem_optin) break;
            if (raise_smem_attr(symbolic_warp_kernel<T>, c.smem) != cudaSuccess) break;
            int ps = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, symbolic_warp_kernel<T>, w * 32, c.smem) != cudaSuccess) break;
            if (ps * w > best_warps) { best_warps = ps * w; best_w = w; best_per_sm = ps; }
        }
        if (best_w > 0) {
            h->wsym = warp_sym_shape<T>((int)kw, (int)pw, (int)now, best_w);
            h->wsym_grid = best_per_sm * h->sm_count;
        }
    }
    // the fast kernel stages a row's previous values and the hot part of its pool block in shared memory, two stages
    // per warp.  Sized for 2.5x the mean block observed so far (the few longer rows are read from global memory), at
    // most for the longest row the symbolic kernel stages, and shrunk until two CTAs fit an SM.
    {
        long long prods = sh.tmp_smem;
        size_t hot = struct_layout<T>(sh.no_max, (int)prods, 1).hot_bytes;
        if (h->avg_hot > 0) {
            size_t want = (size_t)(2.5 * (double)h->avg_hot);
            if (want < 2048) want = 2048;
            if (want < hot) hot = want;
        }
        FastShape f = fast_shape<T>(sh.no_max, (int)hot);
        while (f.smem > (size_t)100 * 1024 && hot > 2048) {
            hot = hot * 3 / 4;
            f = fast_shape<T>(sh.no_max, (int)hot);
        }
        h->fast_no = 0;
        if (f.smem + 1024 <= (size_t)h->max_smem_optin) {
            e = raise_smem_attr(propagate_fast_kernel<T>, f.smem);
            if (e != cudaSuccess) return (int)e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, propagate_fast_kernel<T>, kFastWarps * 32, f.smem);
            if (e != cudaSuccess) return (int)e;
            if (per_sm < 1) per_sm = 1;
            h->fast_no = sh.no_max;
            h->fshape = f;
            h->fast_grid = per_sm * h->sm_count;
            // the big-row kernel: previous values and sums of one row per CTA
            const size_t bs = big_smem_bytes<T>(sh.no_max);
            e = raise_smem_attr(propagate_fast_big_kernel<T>, bs);
            if (e != cudaSuccess) return (int)e;
            int big_per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&big_per_sm, propagate_fast_big_kernel<T>, kBigGroups * 32, bs);
            if (e != cudaSuccess) return (int)e;
            if (big_per_sm < 1) big_per_sm = 1;
            h->big_smem = bs;
            h->big_grid = big_per_sm * h->sm_count;
            if (h->timeline)
                fprintf(stderr, "warp symbolic: kw %d pw %d now %d hw %d rmax %d warps %d per_warp %u smem %zu grid %d\n", h->wsym.kw, h->wsym.pw,
                    h->wsym.now, h->wsym.hw, h->wsym.rmax, h->wsym.warps, h->wsym.per_warp, h->wsym.smem, h->wsym_grid);
            if (h->timeline)
            fprintf(stderr, "shapes: fast no_max %d hot_max %d stage %zu smem %zu CTAs/SM %d (mean hot %lld) | symbolic H %d no_max %d km %d tmp_smem %d smem %zu grid %d\n",
                        f.no_max, f.hot_max, f.stage, f.smem, per_sm, h->avg_hot, sh.H, sh.no_max, sh.km, sh.tmp_smem, sh.smem, h->sym_grid);
        }
    }
    return 0;
}

// Sizes the row-group kernel: table / kept list per row from the caller's hints, lanes per row from the mean
// P row length (one P row per round) and from the number of rows that fit an SM.
template <typename T>
int configure_group_shape(crwr_handle_s* h, long long max_row_len, long long max_kept) {
    const long long n = h->cfg.n_contigs;
    const long long rows = h->cfg.row_end - h->cfg.row_begin;
    long long want = max_row_len + max_row_len / 3 + 32;          // distinct output columns to provide for (rows still grow
                                                                  // by a quarter between the last early poll and the plateau)
    if (want < 48) want = 48;
    if (want > n + 16) want = n + 16;
    // (the longest kept row still grows after the last early poll: +20 % on the dense planted partitions, +30 % on the sparse
    // ones, +50 % in between - 56 -> 84 entries at 220-contig genomes; a kept list that is too short sends the row to the
    // large-row kernel, ~100 us per step, so rows beyond the sparse workloads' 40 entries get the larger margin)
    const long long kept_margin = env_int("CRWR_KEPT_MARGIN_PCT", max_kept > 44 ? 50 : 30);
    long long kmax = round_up(max_kept + max_kept * kept_margin / 100 + 8, 16);
    if (kmax < 32) kmax = 32;
    if (kmax > n) kmax = round_up(n, 16);
    const size_t win_bytes = kHistWin * sizeof(unsigned int);
    const size_t budget = (size_t)h->max_smem_optin - win_bytes;
    // lanes per row: about one P row per round (CRWR_GROUP overrides)
    const double avg_deg = h->p_nnz > 0 ? (double)h->p_nnz / (double)n : 8.0;
    int G = avg_deg <= 6.0 ? 8 : (avg_deg <= 20.0 ? 16 : 32);
    if (const char* e = getenv("CRWR_GROUP")) { const int g = atoi(e); if (g == 4 || g == 8 || g == 16 || g == 32) G = g; }
    const int wmax = 4;
    int wfix = 0;
    if (const char* e = getenv("CRWR_GROUP_WARPS")) { const int g = atoi(e); if (g >= 1 && g <= 8) wfix = g; }
    // overflow list entries = expected longest row / ov_div.  Long rows fill their table to the nominal load more
    // often and spill several products per overflowing column: a quarter instead of an eighth removes the large-row
    // fallbacks of c4_dense (24 -> 1 rows per walk, -10 % per step) and of 220-contig genomes (10 rows in every step,
    // +35 % per walk); short rows keep the smaller list (it would cost them a CTA per SM).
    long long ov_div = want > 320 ? 4 : 8;
    if (const char* e = getenv("CRWR_GROUP_OV_DIV")) { const int v = atoi(e); if (v >= 1 && v <= 64) ov_div = v; }
    // table, kept list and CTA size for one table load
    auto make = [&](int load_pct, long long* in_flight) -> GroupShape {
        GroupShape sh;
        long long w_ = want, km = kmax;
        for (;;) {
            long long H = round_up(w_ * 100 / load_pct + 2, 16);
            if (H > 65520) { H = 65520; w_ = H * load_pct / 100; }
            if (km > H) km = H;
            sh.H = (int)H; sh.limit = (int)round_up(w_, 8); sh.kmax = (int)km;
            sh.ov_cap = (int)(w_ / ov_div < 16 ? 16 : (w_ / ov_div > 512 ? 512 : round_up(w_ / ov_div, 8)));
            sh.per_row = group_smem_per_row<T>(sh.H, sh.limit, sh.kmax, sh.ov_cap);
            if (sh.per_row <= budget) break;
            // too big for one SM: shrink (oversized rows are deferred to the large-row kernel)
            if (km > 1024 && km * 2 > w_) km = round_up(km / 2, 16);
            else w_ = w_ * 3 / 4;
        }
        int g = G;
        while (g < 32 && (size_t)(32 / g) * sh.per_row > budget) g *= 2;
        // warps per CTA: whatever packs the most rows into an SM's shared memory (1 KB reserved per CTA); ties go to
        // the size nearest `wmax`
        const int wcap = (int)(budget / ((size_t)(32 / g) * sh.per_row)) < 8 ? (int)(budget / ((size_t)(32 / g) * sh.per_row)) : 8;
        int w = 1;
        long long best_rows = -1, rows_wmax = -1;
        for (int c = 1; c <= (wcap < 1 ? 1 : wcap); ++c) {
            if (wfix && c != wfix && wfix <= wcap) continue;
            const size_t smem_c = sh.per_row * (size_t)c * (32 / g) + win_bytes;
            long long per_sm_c = (long long)(((size_t)h->max_smem_optin + 1024) / (smem_c + 1024));
            if (per_sm_c * c > 64) per_sm_c = 64 / c;                 // warp slots of an SM
            if (per_sm_c > 32) per_sm_c = 32;
            const long long r = per_sm_c * c;
            const int dist = c > wmax ? c - wmax : wmax - c, dist_w = w > wmax ? w - wmax : wmax - w;
            if (r > best_rows || (r == best_rows && dist < dist_w)) { best_rows = r; w = c; }
            if (c == wmax) rows_wmax = r;
        }
        // a few percent more rows do not pay for the coarser CTAs (measured: 21 vs 20 warps per SM, -0.7 %)
        if (!wfix && rows_wmax > 0 && best_rows * 10 < rows_wmax * 11) w = wmax;
        sh.G = g;
        sh.warps = w;
        const size_t smem = sh.per_row * (size_t)w * (32 / g) + win_bytes;
        const long long per_sm = (long long)(((size_t)h->max_smem_optin + 1024) / (smem + 1024));   // 1 KB reserved per CTA
        *in_flight = per_sm * h->sm_count * w * (32 / g);
        return sh;
    };
    // Table load at the longest expected row (the hint carries a third of growth margin, so real rows sit near 0.5).
    // Measured on B200: 0.70 instead of 0.55 is 10 % faster on long rows and on many rows (more rows per SM); at 0.78
    // the overflow lists of the longest rows start to run out and those rows fall back to the large-row kernel.
    long long fl = 0;
    int base_load = 70;
    if (const char* e = getenv("CRWR_GROUP_LOAD")) { const int v = atoi(e); if (v >= 30 && v <= 85) base_load = v; }
    GroupShape best = make(base_load, &fl);
    GroupShape sh = best;
    sh.rank_sort_max = 128;
    if (const char* e = getenv("CRWR_RANK_SORT_MAX")) sh.rank_sort_max = atoi(e);
    const int g = sh.G, w = sh.warps;
    const size_t smem = sh.per_row * (size_t)w * (32 / g) + win_bytes;
    int per_sm = 0, e = 0;
    switch (g) {
        case 4: e = group_kernel_attr<T, 4>(smem, w * 32, &per_sm, h->max_smem_optin); break;
        case 8: e = group_kernel_attr<T, 8>(smem, w * 32, &per_sm, h->max_smem_optin); break;
        case 16: e = group_kernel_attr<T, 16>(smem, w * 32, &per_sm, h->max_smem_optin); break;
        default: e = group_kernel_attr<T, 32>(smem, w * 32, &per_sm, h->max_smem_optin); break;
    }
    if (e) return e;
    if (per_sm < 1) per_sm = 1;
    h->gshape = sh;
    h->ggrid = per_sm * h->sm_count;
    return 0;
}

unsigned long long* hist_ptr(crwr_handle_s* h);

// Exclusive scan of n int32 counts into n+1 int32 offsets: many CTAs when the input is long.
void launch_scan_i32(crwr_handle_s* h, const int32_t* cnt, int32_t* out, long long n, long long* total_out, cudaStream_t st) {
    const long long tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles >= 2 && tiles <= kScanMaxTiles) {
        long long* tmp = h->at<long long>(h->L.scan_tmp);
        scan_tile_sums_kernel<<<(int)tiles, 1024, 0, st>>>(cnt, n, tmp);
        scan_tiles_kernel<int32_t><<<(int)tiles, 1024, 0, st>>>(cnt, out, n, tmp, total_out);
        g_launches += 2;
    } else {
        exclusive_scan_kernel<int32_t><<<1, 1024, 0, st>>>(cnt, out, n, total_out);
        ++g_launches;
    }
}

// Programmatic dependent launch of the step kernels (common.cuh, pdl_wait): hides the launch gaps between the four
// dependent kernels of a step.  Measured on B200: -9 % per step at ~100-entry rows (90 -> 81 us), +4 % at ~650-entry
// rows (673 -> 700 us; CTAs parked in pdl_wait share the SMs with long-running predecessors), so it is used for
// short rows only.  CRWR_PDL=0/1 forces it off/on.
bool use_pdl(const crwr_handle_s* h) {
    // (sharded walks: only the peer-attached chain - its barrier kernels wait for their predecessor like everybody
    // else; the host-driven collectives of the other path sit between the kernels anyway)
    if (h->cfg.n_shards != 1 && !h->peer_attached) return false;
    if (h->pdl_forced >= 0) return h->pdl_forced != 0;
    return h->hint_len_eff <= 384;
}

template <typename T>
PropArgs<T> prop_args(crwr_handle_s* h, int parity, bool fused) {
    const Layout& L = h->L;
    PropArgs<T> a;
    a.p_indptr = h->at<int32_t>(L.p_indptr);
    a.p_cols = h->at<int32_t>(L.p_cols);
    a.p_vals = h->at<T>(L.p_vals);
    const int in = parity, out = parity ^ 1;
    a.in = IterBuf<T>{h->at<int32_t>(L.it_cols[in]), h->at<T>(L.it_vals[in]), h->at<long long>(L.it_ptr[in]),
                      h->at<int32_t>(L.it_len[in])};
    a.out = IterBuf<T>{h->at<int32_t>(L.it_cols[out]), h->at<T>(L.it_vals[out]), h->at<long long>(L.it_ptr[out]),
                       h->at<int32_t>(L.it_len[out])};
    a.out_cap = h->cfg.iterate_cap;
    a.st = h->at<WalkState>(L.state);
    a.fb_rows = h->at<int32_t>(L.fb_rows);
    a.n_rows = (int32_t)(h->cfg.row_end - h->cfg.row_begin);
    a.row_begin = (int32_t)h->cfg.row_begin;
    a.scale = (T)(1.0 - h->rp);                 // (1 - rp) * matrix: the scalar takes the matrix dtype
    a.restart = (T)(float)h->rp;                // rp * I with I float32 (utility.py:278)
    a.hist = fused ? hist_ptr(h) : nullptr;
    a.fused_select = fused ? 1 : 0;
    a.prefill_l1 = fused ? 1 : 0;
    a.row_order = h->order_valid ? h->at<int32_t>(L.row_order) : nullptr;
    a.row_list_count = nullptr;
    a.rs = h->at<RowStruct>(L.rs);
    a.pool = h->at<unsigned char>(L.pool);
    a.pool_cap = (unsigned long long)L.pool_bytes;
    a.slow_list = h->at<int32_t>(L.slow_list);
    a.alpha = h->steps_enqueued == h->build_step ? h->alpha_build : h->alpha;
    a.fixed_base = L.fixed_base;
    a.slot_cap = L.slot_cap;
    a.sym_tmp = h->at<unsigned char>(L.sym_tmp);
    a.small_no = h->fshape.no_max;
    a.small_hot = h->fshape.hot_max;

    {
        // while many rows are built a warp reserves slot entries / pool bytes for about four rows at a time
        long long sc = 4 * (long long)h->hint_len_eff;
        if (sc < 256) sc = 256;
        if (sc > 65536) sc = 65536;
        a.slot_chunk = (int32_t)sc;
        a.pool_chunk = (int32_t)(sc * 32 < (1 << 20) ? sc * 32 : (1 << 20));
    }
    a.win_mode = 0;
    a.cand = nullptr;
    a.cand_cap = (unsigned long long)L.cand_cap;
    // a warp reserves output space for a few rows at a time (zero-padded leftovers are skipped by the select)
    {
        long long ch = 3 * (long long)h->hint_len_eff;
        if (ch < 256) ch = 256;
        if (ch > 8192) ch = 8192;
        a.chunk = (int32_t)ch;
        // rows per scheduling request: one while every warp gets at most a few rows, more when there are many
        const long long warps = (long long)h->grid * h->shape.warps;
        long long rc = a.n_rows / (warps * 2 > 0 ? warps * 2 : 1);
        rc = 1;      // measured: one row per request is fastest at every size (2 and 4 lose to imbalance)
        if (const char* e = getenv("CRWR_ROW_CHUNK")) rc = atoi(e) > 0 ? atoi(e) : rc;
        a.row_chunk = (int32_t)rc;
    }
    return a;
}

unsigned long long* cand_ptr(crwr_handle_s* h, int step);
void tl_mark(crwr_handle_s* h, cudaStream_t st, const char* label);

template <typename T>
void launch_group(const GroupShape& sh, bool pdl, int grid, size_t smem, cudaStream_t st, const PropArgs<T>& a) {
    switch (sh.G) {
        case 4: launch_k(pdl, propagate_group_kernel<T, 4>, grid, sh.warps * 32, smem, st, a, sh); break;
        case 8: launch_k(pdl, propagate_group_kernel<T, 8>, grid, sh.warps * 32, smem, st, a, sh); break;
        case 16: launch_k(pdl, propagate_group_kernel<T, 16>, grid, sh.warps * 32, smem, st, a, sh); break;
        default: launch_k(pdl, propagate_group_kernel<T, 32>, grid, sh.warps * 32, smem, st, a, sh); break;
    }
}

// The propagation of one walk step.  Step 0: warp-per-row kernel (keeps scipy's storage order).  Steps before
// build_step: row-group kernel.  From build_step on: [fast kernel: rows with a current cached structure] -> symbolic
// kernel (the rest: structure build + numeric pass).  Behind either: [large-row kernel: deferred rows].
// fused: the large-row kernel's last CTA scans the histograms (unsharded walk, histogram select).
// win_step: the step's cut comes from the window select (tail.cuh): values are classified against the window.
template <typename T>
int launch_propagate(crwr_handle_s* h, int parity, bool fused, bool win_step, cudaStream_t st) {
    PropArgs<T> a = prop_args<T>(h, parity, fused);
    if (!fused && h->cfg.n_shards > 1 && h->hist0_fused) {
        a.hist = hist_ptr(h);                       // sharded: histogram only, no scan tail
        a.prefill_l1 = h->peer_attached ? 1 : 0;    // peer exchange: level 1 pre-filled for the predicted bin
    }
    const int step = h->steps_enqueued;
    if (win_step) {
        a.hist = hist_ptr(h);                       // used when the previous step could not set a window
        a.prefill_l1 = 1;
        a.fused_select = 0;
        a.win_mode = 1;
        a.cand = cand_ptr(h, step);
    }
    const int rows = a.n_rows;
    const bool pdl = use_pdl(h);
    // (a sharded walk may have to run a step twice - peer.cuh, window miss - which a freshly rebuilt structure cannot)
    if (step == h->build_step)
        h->structs_active = h->cfg.n_shards == 1 &&
                            (h->struct_mode == 2 || (h->struct_mode == 1 && h->mean_row >= h->struct_min_row));
    const bool structs = h->structs_active && h->prop_kind == 1 && step >= h->build_step && h->fast_no > 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profile) {
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
        cudaEventRecord(e0, st);
    }
    tl_mark(h, st, "step start");
    if (structs) {
        a.row_order = nullptr;
        if (step > h->build_step && h->big_seen) {
            launch_k(pdl, propagate_fast_big_kernel<T>, rows < h->big_grid ? rows : h->big_grid, kBigGroups * 32, h->big_smem, st, a,
                     h->fshape, h->sym.no_max);
            ++g_launches;
            if (cudaGetLastError() != cudaSuccess) return -1;
            tl_mark(h, st, "fast (long rows)");
        }
        if (step > h->build_step) {
            int grid = (rows + kFastWarps - 1) / kFastWarps;
            if (grid > h->fast_grid) grid = h->fast_grid;
            launch_k(pdl, propagate_fast_kernel<T>, grid, kFastWarps * 32, h->fshape.smem, st, a, h->fshape);
            ++g_launches;
            if (cudaGetLastError() != cudaSuccess) return -1;
            tl_mark(h, st, "fast");
            // the symbolic kernel walks the rows the fast kernel handed over
            a.row_order = h->at<int32_t>(h->L.slow_list);
            a.row_list_count = &h->at<WalkState>(h->L.state)->slow_count;
        }
        if (h->wsym.warps > 0 && step == h->build_step) {
            // the first build of short rows: a warp each (a single row is built sooner by a CTA, so later lists go straight
            // to the CTA kernel); what exceeds a warp's tables comes back through the second list
            int32_t* list2 = h->at<int32_t>(h->L.slow_list2);
            int wgrid = (rows + h->wsym.warps - 1) / h->wsym.warps;
            if (wgrid > h->wsym_grid) wgrid = h->wsym_grid;
            launch_k(pdl, symbolic_warp_kernel<T>, wgrid, h->wsym.warps * 32, h->wsym.smem, st, a, h->wsym, list2);
            ++g_launches;
            if (cudaGetLastError() != cudaSuccess) return -1;
            tl_mark(h, st, "symbolic (warp)");
            a.row_order = list2;
            a.row_list_count = &h->at<WalkState>(h->L.state)->slow2_count;
        }
        const SymShape& sh = h->sym;
        int grid = rows < h->sym_grid ? rows : h->sym_grid;
        // the first build has a row for every CTA (smaller CTAs, more rows in flight); later lists are short (larger
        // CTAs, shorter latency per row)
        launch_k(pdl, symbolic_kernel<T>, grid, step > h->build_step ? sh.threads : sh.threads_build, sh.smem, st, a, sh);
    } else if (h->prop_kind == 1 && step >= 1) {      // step 0 keeps scipy's storage order (warp-per-row kernel)
        if (h->use_order && !h->order_valid && step >= kOrderStep && rows >= 64) {
            // the walk has settled (the host sized the kernels from the third poll): rows by length, once
            row_order_kernel<<<1, 1024, 0, st>>>(h->at<int32_t>(h->L.it_len[parity]), rows, h->at<int32_t>(h->L.row_order));
            ++g_launches;
            h->order_valid = true;
            a.row_order = h->at<int32_t>(h->L.row_order);
        }
        const GroupShape& sh = h->gshape;
        const int rpc = sh.warps * (32 / sh.G);          // rows per CTA and pass
        int grid = (rows + rpc - 1) / rpc;
        if (grid > h->ggrid) grid = h->ggrid;
        if (grid < 1) grid = 1;
        const size_t smem = sh.per_row * (size_t)rpc + kHistWin * sizeof(unsigned int);
        launch_group<T>(sh, pdl, grid, smem, st, a);
    } else {
        const PropShape& sh = h->shape;
        int grid = (rows + sh.warps - 1) / sh.warps;
        if (grid > h->grid) grid = h->grid;
        if (grid < 1) grid = 1;
        launch_k(pdl, propagate_kernel<T>, grid, sh.warps * 32, sh.per_warp * (size_t)sh.warps + kHistWin * sizeof(unsigned int), st, a, sh);
    }
    ++g_launches;
    tl_mark(h, st, "propagate/symbolic");
    if (h->profile) {
        cudaEventRecord(e1, st);
        h->ev.push_back(e0);
        h->ev.push_back(e1);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    // the large-row kernel: always behind a histogram-select step (its last CTA scans the histograms); behind a
    // window-select step only when this walk has deferred rows before (a deferral it then misses is caught by the finish
    // step and the host retries with the kernel always queued)
    // (a sharded histogram step needs no scan tail either: same rule as a window step)
    h->dense_launched = (!win_step && (fused || h->cfg.n_shards == 1)) || h->dense_always || h->fb_seen;
    if (h->dense_launched) {
        a.row_order = nullptr;
        a.row_list_count = nullptr;
        const Layout& L = h->L;
        DenseScratch<T> sc;
        sc.acc = h->at<T>(L.d_acc);
        sc.oldv = h->at<T>(L.d_oldv);
        sc.ord = h->at<int32_t>(L.d_ord);
        sc.tbits = h->at<uint32_t>(L.d_tbits);
        sc.kbits = h->at<uint32_t>(L.d_kbits);
        sc.n = h->cfg.n_contigs;
        sc.words = L.words;
        launch_k(pdl, propagate_dense_kernel<T, kDenseWarpsPerCta>, (int)(L.dense_warps / kDenseWarpsPerCta), kDenseWarpsPerCta * 32, 0, st, a, sc);
        ++g_launches;
        if (cudaGetLastError() != cudaSuccess) return -1;
        tl_mark(h, st, "large rows");
    }
    return 0;
}

// Where this rank's step contributions go: its symmetric block when the peer exchange is attached (peers
// read it over NVLink), else the local exchange buffer (which then also receives the all-reduced sums).
unsigned long long* xbuf_ptr(crwr_handle_s* h) {
    if (h->peer_attached) return h->peer_block + kPeerX;
    return h->at<unsigned long long>(h->L.xbuf);
}
unsigned long long* hist_ptr(crwr_handle_s* h) { return xbuf_ptr(h) + kXExtras; }
unsigned long long* cand_ptr(crwr_handle_s* h, int step) {
    if (h->peer_attached) return h->peer_block + kPeerCand0 + (step & 1) * kPeerCandWords;
    return h->at<unsigned long long>(h->L.cand);
}

template <typename T>
int launch_hist(crwr_handle_s* h, int out_parity, int level, int fused, cudaStream_t st) {
    const Layout& L = h->L;
    launch_k(use_pdl(h), select_hist_kernel<T>, h->sm_count, 512, 0, st, (const T*)h->at<T>(L.it_vals[out_parity]),
             L.fixed_base, h->at<WalkState>(L.state), hist_ptr(h), level, fused);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <typename T>
int launch_scan(crwr_handle_s* h, int level, cudaStream_t st) {
    const Layout& L = h->L;
    select_scan_kernel<T><<<1, 1024, 0, st>>>(h->at<WalkState>(L.state), hist_ptr(h), level);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
FinishArgs finish_args(crwr_handle_s* h, int do_select, const void* gathered) {
    const Layout& L = h->L;
    FinishArgs a;
    a.st = h->at<WalkState>(L.state);
    a.hist = h->at<unsigned long long>(L.xbuf);
    a.xch = h->cfg.n_shards > 1 ? h->at<unsigned long long>(L.xbuf) : nullptr;
    a.own_block = h->at<unsigned long long>(L.cand);
    a.blocks = gathered ? (unsigned long long*)gathered : a.own_block;
    a.n_blocks = gathered ? h->cfg.n_shards : 1;
    a.block_stride = kCandHeader + L.cand_cap;
    a.merged = h->at<unsigned long long>(L.merged);
    a.cand_cap = (unsigned long long)L.cand_cap;
    a.trace = h->at<crwr_step_record>(L.trace);
    a.do_select = do_select;
    // the step after this one is a window-select step: derive its window from this step's cut
    a.next_win = (h->win_step > 0 && (h->cfg.n_shards == 1 || h->peer_attached) && h->steps_enqueued + 1 >= h->win_step &&
                  h->steps_enqueued + 1 >= h->win_hold_until) ? 1 : 0;
    a.dense_launched = h->dense_launched ? 1 : 0;
    a.win_scale = h->win_scale; a.win_min = h->win_min; a.win_max = h->win_max; a.win_rel_max = h->win_rel_max;
    a.win_narrow_cands = (long long)h->cfg.n_shards * (kPushCand * 3 / 4);
    return a;
}

template <typename T>
int launch_gather(crwr_handle_s* h, int out_parity, int fused, cudaStream_t st) {
    const Layout& L = h->L;
    launch_k(use_pdl(h), select_gather_kernel<T>, h->sm_count, 512, 0, st, (const T*)h->at<T>(L.it_vals[out_parity]),
             L.fixed_base, h->at<WalkState>(L.state), cand_ptr(h, h->steps_enqueued), (unsigned long long)L.cand_cap, fused,
             finish_args(h, 1, nullptr));
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
// gathered: candidate blocks of all ranks (sharded) or null (own block only)
template <typename T>
int launch_finish(crwr_handle_s* h, int do_select, const void* gathered, cudaStream_t st) {
    launch_k(use_pdl(h), finish_step_kernel<T>, 1, 1024, 0, st, finish_args(h, do_select, gathered));
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
int launch_tail(crwr_handle_s* h, int out_parity, cudaStream_t st) {
    const Layout& L = h->L;
    TailArgs a;
    a.fin = finish_args(h, 1, nullptr);
    a.vals = h->at<T>(L.it_vals[out_parity]);
    a.fixed_base = L.fixed_base;
    a.cand = cand_ptr(h, h->steps_enqueued);
    a.cand_cap = (unsigned long long)L.cand_cap;
    launch_k(use_pdl(h), tail_kernel<T>, 1, 1024, 0, st, a);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
DenseScratch<T> dense_scratch(crwr_handle_s* h) {
    const Layout& L = h->L;
    DenseScratch<T> sc;
    sc.acc = h->at<T>(L.d_acc);
    sc.oldv = h->at<T>(L.d_oldv);
    sc.ord = h->at<int32_t>(L.d_ord);
    sc.tbits = h->at<uint32_t>(L.d_tbits);
    sc.kbits = h->at<uint32_t>(L.d_kbits);
    sc.n = h->cfg.n_contigs;
    sc.words = L.words;
    return sc;
}

void tl_mark(crwr_handle_s* h, cudaStream_t st, const char* label) {
    if (!h->profile || !h->timeline) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    h->tl.push_back(std::make_pair(e, label));
}

void prof_mark(crwr_handle_s* h, cudaStream_t st, int id) {
    if (!h->profile) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    h->marks.push_back(e);
    h->mark_ids.push_back(id);
}

template <typename T>
int enqueue_step(crwr_handle_s* h, cudaStream_t st) {
    // the step index and buffer parity are deterministic on the host: every enqueued step either
    // runs completely or (after the stop rule fired) is skipped completely.
    const int step = h->steps_enqueued;
    const int parity = step & 1;
    if (step == 0) {
        // no percentile cut on step 0 (utility.py:286)
        if (launch_propagate<T>(h, parity, false, false, st)) return -1;
        if (launch_finish<T>(h, 0, nullptr, st)) return -1;
    } else if (h->win_step > 0 && step >= h->win_step && step >= h->win_hold_until) {
        // settled walk: propagation kernels (values classified against the select window) -> one single-CTA tail
        prof_mark(h, st, 0);
        if (launch_propagate<T>(h, parity, false, true, st)) return -1;
        prof_mark(h, st, 1);
        prof_mark(h, st, 2);
        if (launch_tail<T>(h, parity ^ 1, st)) return -1;
        tl_mark(h, st, "tail");
        prof_mark(h, st, 3);
    } else {
        // propagate (+ level-0/1 histograms in its epilogue) -> large-row kernel (+ scans in its last CTA)
        // -> level-1 histogram pass (skipped on the device when the prediction held) -> gather (+ finish)
        prof_mark(h, st, 0);
        if (launch_propagate<T>(h, parity, true, false, st)) return -1;
        prof_mark(h, st, 1);
        if (launch_hist<T>(h, parity ^ 1, 1, 1, st)) return -1;
        tl_mark(h, st, "hist1");
        prof_mark(h, st, 2);
        if (launch_gather<T>(h, parity ^ 1, 1, st)) return -1;
        tl_mark(h, st, "gather+finish");
        prof_mark(h, st, 3);
    }
    h->steps_enqueued = step + 1;
    return 0;
}

template <typename T>
int count_rows(crwr_handle_s* h, int col_limit, int64_t* h_nnz, cudaStream_t st) {
    const Layout& L = h->L;
    const int rows = (int)(h->cfg.row_end - h->cfg.row_begin);
    WalkState host_state;
    if (cudaMemcpyAsync(&host_state, h->ws + L.state, sizeof(WalkState), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    const int parity = host_state.parity;
    IterBuf<T> buf{h->at<int32_t>(L.it_cols[parity]), h->at<T>(L.it_vals[parity]), h->at<long long>(L.it_ptr[parity]),
                   h->at<int32_t>(L.it_len[parity])};
    const int blocks = (rows * 32 + 255) / 256;
    result_count_kernel<T><<<blocks > 0 ? blocks : 1, 256, 0, st>>>(buf, h->at<WalkState>(L.state), rows, col_limit,
                                                                   h->at<int32_t>(L.counts));
    ++g_launches;
    exclusive_scan_kernel<int64_t><<<1, 1024, 0, st>>>(h->at<int32_t>(L.counts), h->at<int64_t>(L.counts_scan), rows,
                                                       h->at<long long>(L.total));
    ++g_launches;
    long long total = 0;
    if (cudaMemcpyAsync(&total, h->ws + L.total, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (cudaGetLastError() != cudaSuccess) return -1;
    *h_nnz = total;
    return 0;
}

template <typename T>
int write_rows(crwr_handle_s* h, int col_limit, int64_t* d_indptr, int32_t* d_indices, void* d_data, cudaStream_t st) {
    const Layout& L = h->L;
    const int rows = (int)(h->cfg.row_end - h->cfg.row_begin);
    WalkState host_state;
    if (cudaMemcpyAsync(&host_state, h->ws + L.state, sizeof(WalkState), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    const int parity = host_state.parity;
    IterBuf<T> buf{h->at<int32_t>(L.it_cols[parity]), h->at<T>(L.it_vals[parity]), h->at<long long>(L.it_ptr[parity]),
                   h->at<int32_t>(L.it_len[parity])};
    if (cudaMemcpyAsync(d_indptr, h->ws + L.counts_scan, 8 * (size_t)(rows + 1), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return -1;
    const int blocks = (rows * 32 + 255) / 256;
    result_write_kernel<T><<<blocks > 0 ? blocks : 1, 256, 0, st>>>(buf, h->at<WalkState>(L.state), rows, col_limit,
                                                                   h->at<int64_t>(L.counts_scan), d_indices, (T*)d_data);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" {

int32_t crwr_abi_version(void) { return CRWR_ABI_VERSION; }
const char* crwr_last_error(void) { return g_err.c_str(); }
int64_t crwr_launch_count(int32_t reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

static int check_cfg(const crwr_config* c) {
    if (!c) CRWR_FAIL(CRWR_ERR_INVALID, "null config");
    if (c->dtype != CRWR_F32 && c->dtype != CRWR_F64) CRWR_FAIL(CRWR_ERR_INVALID, "dtype must be CRWR_F32 or CRWR_F64");
    if (c->n_contigs < 1 || c->n_contigs > 2000000000ll) CRWR_FAIL(CRWR_ERR_INVALID, "n_contigs out of range");
    if (c->n_markers < 1 || c->n_markers > c->n_contigs) CRWR_FAIL(CRWR_ERR_INVALID, "n_markers must be in [1, n_contigs]");
    if (c->row_begin < 0 || c->row_end > c->n_markers || c->row_begin >= c->row_end)
        CRWR_FAIL(CRWR_ERR_INVALID, "row range must be a non-empty sub-range of [0, n_markers)");
    if (c->transition_nnz_cap < 1 || c->transition_nnz_cap > 2000000000ll) CRWR_FAIL(CRWR_ERR_INVALID, "transition_nnz_cap out of range");
    if (c->iterate_cap < c->row_end - c->row_begin) CRWR_FAIL(CRWR_ERR_INVALID, "iterate_cap smaller than the row count");
    if (c->n_shards < 1 || c->n_shards > 1024) CRWR_FAIL(CRWR_ERR_INVALID, "n_shards must be in [1, 1024]");
    if (c->n_shards == 1 && (c->row_begin != 0 || c->row_end != c->n_markers))
        CRWR_FAIL(CRWR_ERR_INVALID, "an unsharded walk owns all restart rows");
    return CRWR_OK;
}

int64_t crwr_workspace_bytes(const crwr_config* cfg) {
    if (check_cfg(cfg) != CRWR_OK) return -1;
    Layout L;
    make_layout(*cfg, L);
    return (int64_t)L.bytes;
}

int32_t crwr_create(crwr_handle* out, const crwr_config* cfg, void* d_workspace, int64_t workspace_bytes) {
    if (!out) CRWR_FAIL(CRWR_ERR_INVALID, "null out");
    int rc = check_cfg(cfg);
    if (rc != CRWR_OK) return rc;
    if (!d_workspace) CRWR_FAIL(CRWR_ERR_INVALID, "null workspace");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) CRWR_FAIL(CRWR_ERR_CUDA, "no CUDA device: libcrwr_b200 has no CPU path");
    CUDA_TRY(cudaSetDevice(cfg->device));
    crwr_handle_s* h = new crwr_handle_s();
    h->cfg = *cfg;
    make_layout(*cfg, h->L);
    if ((int64_t)h->L.bytes > workspace_bytes) {
        const long long need = (long long)h->L.bytes;
        delete h;
        CRWR_FAIL(CRWR_ERR_INVALID, "workspace too small: need %lld bytes", need);
    }
    h->ws = (unsigned char*)d_workspace;
    int sm_count = 0, smem_optin = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    CUDA_TRY(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    h->sm_count = sm_count;
    h->p_nnz = 0;
    h->have_p = false;
    h->walking = false;
    h->steps_enqueued = 0;
    h->max_smem_optin = smem_optin;
    h->h_state = acquire_pinned_state();
    if (!h->h_state) {
        delete h;
        CRWR_FAIL(CRWR_ERR_CUDA, "cudaMallocHost failed");
    }
    h->bitmaps_zeroed = false;
    // the large-row kernel's bitmaps clean up after themselves: they are zeroed once, on the caller's stream, by the
    // first crwr_walk_init (bitmaps_zeroed)
    int e = cfg->dtype == CRWR_F64 ? configure_shape<double>(h, 192, 32) : configure_shape<float>(h, 192, 32);
    if (!e) e = cfg->dtype == CRWR_F64 ? configure_group_shape<double>(h, 192, 32) : configure_group_shape<float>(h, 192, 32);
    if (!e) e = cfg->dtype == CRWR_F64 ? configure_struct_shape<double>(h, 192, 32) : configure_struct_shape<float>(h, 192, 32);
    if (e) {
        release_pinned_state(h->h_state);
        delete h;
        CRWR_FAIL(CRWR_ERR_CUDA, "kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)e));
    }
    if (const char* k = getenv("CRWR_PROP")) h->prop_kind = strcmp(k, "warp") == 0 ? 0 : 1;
    if (const char* k = getenv("CRWR_PDL")) h->pdl_forced = atoi(k) != 0 ? 1 : 0;
    if (const char* k = getenv("CRWR_ROW_ORDER")) h->use_order = atoi(k) != 0;
    h->struct_mode = env_int("CRWR_STRUCT", 1);
    if (h->struct_mode < 0 || h->struct_mode > 2) h->struct_mode = 1;
    h->struct_min_row = env_int("CRWR_STRUCT_MIN_ROW", 176);
    h->timeline = env_int("CRWR_TIMELINE", 0) != 0;
    h->warp_sym = env_int("CRWR_WARP_SYM", 1);
    h->build_step = env_int("CRWR_BUILD_STEP", 3);
    if (h->build_step < 2) h->build_step = 2;       // structures only exist behind a cut
    h->win_step = env_int("CRWR_WIN_STEP", 6);
    if (h->win_step > 0 && h->win_step < 2) h->win_step = 2;
    h->win_step_sharded = env_int("CRWR_WIN_STEP_SHARDED", 9);
    h->scan_push = env_int("CRWR_SCAN_PUSH", 1) != 0;

    if (h->prop_kind != 1) h->win_step = 0;         // the warp-per-row kernel only knows the histogram select
    h->alpha = (float)env_double("CRWR_STRUCT_ALPHA", 0.5);
    if (!(h->alpha > 0.f && h->alpha <= 1.f)) h->alpha = 0.5f;
    h->alpha_build = (float)env_double("CRWR_STRUCT_ALPHA_BUILD", 0.8 * (double)h->alpha);
    if (!(h->alpha_build > 0.f && h->alpha_build <= 1.f)) h->alpha_build = h->alpha;
    h->win_scale = (float)env_double("CRWR_WIN_SCALE", 2.5);
    h->win_min = (float)env_double("CRWR_WIN_MIN", 0.004);
    h->win_max = (float)env_double("CRWR_WIN_MAX", 0.15);
    h->win_rel_max = (float)env_double("CRWR_WIN_REL_MAX", 0.10);
    *out = h;
    return CRWR_OK;
}

int32_t crwr_destroy(crwr_handle h) {
    if (h) {
        for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
        if (h->h_state) release_pinned_state(h->h_state);
    }
    delete h;
    return CRWR_OK;
}

int32_t crwr_build_transition(crwr_handle h, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                              int64_t nnz, const int32_t* d_new_of_old, int64_t* h_nnz_out, void* stream) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = h->L;
    const long long n = h->cfg.n_contigs;
    if (nnz < 0 || nnz + n > h->cfg.transition_nnz_cap)
        CRWR_FAIL(CRWR_ERR_INVALID, "transition_nnz_cap must be at least nnz + n_contigs (self-loops)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    TransitionScratch sc;
    sc.row_cnt = h->at<int32_t>(L.k1_row_cnt);
    sc.col_cnt = h->at<int32_t>(L.k1_col_cnt);
    sc.col_ptr = h->at<int32_t>(L.k1_col_ptr);
    sc.col_fill = h->at<int32_t>(L.k1_col_fill);
    sc.csc_row = h->at<int32_t>(L.k1_csc_row);
    sc.csc_val = h->at<double>(L.k1_csc_val);
    sc.scale = h->at<double>(L.k1_scale);
    sc.isolated = h->at<int32_t>(L.k1_iso);
    sc.row_val = h->at<double>(L.k1_row_val);
    sc.total = h->at<long long>(L.total);
    sc.bad = h->at<int32_t>(L.k1_bad);
    CUDA_TRY(cudaMemsetAsync(sc.bad, 0, 8, st));
    CUDA_TRY(cudaMemsetAsync(sc.col_cnt, 0, 4 * (size_t)(n + 1), st));
    CUDA_TRY(cudaMemsetAsync(sc.col_fill, 0, 4 * (size_t)n, st));
    CUDA_TRY(cudaMemsetAsync(sc.row_cnt, 0, 4 * (size_t)(n + 1), st));
    const int tb = 256;
    const int rows_grid = (int)((n + tb - 1) / tb);
    const int warp_grid = (int)((n * 32 + tb - 1) / tb);
    transition_count_kernel<<<rows_grid, tb, 0, st>>>(d_indptr, d_indices, d_data, d_new_of_old, n, sc);
    LAUNCH_CHECK();
    launch_scan_i32(h, sc.col_cnt, sc.col_ptr, n, nullptr, st);
    transition_scatter_kernel<<<rows_grid, tb, 0, st>>>(d_indptr, d_indices, d_data, d_new_of_old, n, sc);
    LAUNCH_CHECK();
    transition_colsum_kernel<<<warp_grid, tb, 0, st>>>(n, sc);
    LAUNCH_CHECK();
    launch_scan_i32(h, sc.row_cnt, h->at<int32_t>(L.p_indptr), n, sc.total, st);
    if (h->cfg.dtype == CRWR_F64)
        transition_fill_kernel<double><<<warp_grid, tb, 0, st>>>(d_indptr, d_indices, d_data, d_new_of_old, n, sc,
                                                                 h->at<int32_t>(L.p_indptr), h->at<int32_t>(L.p_cols),
                                                                 h->at<double>(L.p_vals));
    else
        transition_fill_kernel<float><<<warp_grid, tb, 0, st>>>(d_indptr, d_indices, d_data, d_new_of_old, n, sc,
                                                                h->at<int32_t>(L.p_indptr), h->at<int32_t>(L.p_cols),
                                                                h->at<float>(L.p_vals));
    LAUNCH_CHECK();
    long long total = 0;
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, sc.total, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&bad, sc.bad, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (bad) CRWR_FAIL(CRWR_ERR_INVALID, "contact values must be finite and non-negative");
    h->p_nnz = total;
    h->have_p = true;

    if (h_nnz_out) *h_nnz_out = total;
    return CRWR_OK;
}

int32_t crwr_set_transition(crwr_handle h, const int32_t* d_indptr, const int32_t* d_indices, const void* d_data,
                            int64_t nnz, void* stream) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    if (nnz < 0 || nnz > h->cfg.transition_nnz_cap) CRWR_FAIL(CRWR_ERR_INVALID, "nnz exceeds transition_nnz_cap");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = h->L;
    const size_t s = elem(h->cfg.dtype);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(h->ws + L.p_indptr, d_indptr, 4 * (size_t)(h->cfg.n_contigs + 1), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->ws + L.p_cols, d_indices, 4 * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->ws + L.p_vals, d_data, s * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
    h->p_nnz = nnz;
    h->have_p = true;

    return CRWR_OK;
}

int32_t crwr_get_transition(crwr_handle h, int32_t* d_indptr, int32_t* d_indices, void* d_data, void* stream) {
    if (!h || !h->have_p) CRWR_FAIL(CRWR_ERR_STATE, "no transition matrix");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = h->L;
    const size_t s = elem(h->cfg.dtype);
    CUDA_TRY(cudaMemcpyAsync(d_indptr, h->ws + L.p_indptr, 4 * (size_t)(h->cfg.n_contigs + 1), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_indices, h->ws + L.p_cols, 4 * (size_t)h->p_nnz, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_data, h->ws + L.p_vals, s * (size_t)h->p_nnz, cudaMemcpyDeviceToDevice, st));
    return CRWR_OK;
}

int64_t crwr_transition_nnz(crwr_handle h) { return h && h->have_p ? h->p_nnz : -1; }

int32_t crwr_walk_init(crwr_handle h, double rp, double perc, double tol, int32_t max_steps, void* stream) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    if (!h->have_p) CRWR_FAIL(CRWR_ERR_STATE, "transition matrix not set");
    if (!(rp >= 0.0 && rp <= 1.0)) CRWR_FAIL(CRWR_ERR_INVALID, "rp must be in [0,1]");
    if (!(perc >= 0.0 && perc <= 100.0)) CRWR_FAIL(CRWR_ERR_INVALID, "Percentiles must be in the range [0, 100]");
    if (max_steps < 1 || max_steps > kMaxTrace) CRWR_FAIL(CRWR_ERR_INVALID, "max_steps must be in [1, %d]", kMaxTrace);
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = h->L;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (!h->bitmaps_zeroed) {
        CUDA_TRY(cudaMemsetAsync(h->ws + L.d_tbits, 0, 4 * (size_t)L.words * L.dense_warps, st));
        CUDA_TRY(cudaMemsetAsync(h->ws + L.d_kbits, 0, 4 * (size_t)L.words * L.dense_warps, st));
        h->bitmaps_zeroed = true;
    }
    init_state_kernel<<<1, 256, 0, st>>>(h->at<WalkState>(L.state), rp, perc, tol, max_steps,
                                         h->at<unsigned long long>(L.xbuf), kXExtras + kSelectLevels * kSelectBins,
                                         h->at<unsigned long long>(L.cand));
    LAUNCH_CHECK();
    const int rows = (int)(h->cfg.row_end - h->cfg.row_begin);
    if (h->cfg.dtype == CRWR_F64) {
        IterBuf<double> b{h->at<int32_t>(L.it_cols[0]), h->at<double>(L.it_vals[0]), h->at<long long>(L.it_ptr[0]),
                          h->at<int32_t>(L.it_len[0])};
        init_iterate_kernel<double><<<(rows + 255) / 256, 256, 0, st>>>(b, rows, (int)h->cfg.row_begin);
    } else {
        IterBuf<float> b{h->at<int32_t>(L.it_cols[0]), h->at<float>(L.it_vals[0]), h->at<long long>(L.it_ptr[0]),
                         h->at<int32_t>(L.it_len[0])};
        init_iterate_kernel<float><<<(rows + 255) / 256, 256, 0, st>>>(b, rows, (int)h->cfg.row_begin);
    }
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemsetAsync(h->ws + L.rs, 0, sizeof(RowStruct) * (size_t)rows, st));     // generation 0: no structure
    h->rp = rp;
    h->perc = perc;
    h->steps_enqueued = 0;
    h->win_hold_until = 0;
    h->avg_hot = 0;
    h->mean_row = 0;
    h->structs_active = false;
    h->sharded_window_ok = false;
    h->redo_hist = false;
    h->big_seen = false;
    h->fb_seen = false;
    h->dense_launched = true;
    h->order_valid = false;
    h->walking = true;
    return CRWR_OK;
}

int64_t crwr_bin_contact_scratch_bytes(int64_t m) { return m < 1 ? -1 : 8 * m; }

int32_t crwr_bin_contact_weights(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices, const void* d_data,
                                 int64_t n_rows, const int32_t* d_rows, int64_t n_bins, const int64_t* d_bin_ptr,
                                 const int32_t* d_bin_members, int64_t n_members, void* d_scratch, int64_t scratch_bytes,
                                 void* d_out, void* stream) {
    if (dtype != CRWR_F32 && dtype != CRWR_F64) CRWR_FAIL(CRWR_ERR_INVALID, "dtype must be CRWR_F32 or CRWR_F64");
    if (m < 1 || n_rows < 0 || n_bins < 0 || n_members < 0) CRWR_FAIL(CRWR_ERR_INVALID, "bad sizes");
    if (n_rows == 0 || n_bins == 0) return CRWR_OK;
    if (!d_indptr || !d_indices || !d_data || !d_rows || !d_bin_ptr || !d_bin_members || !d_out)
        CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    if (d_scratch && scratch_bytes < 8 * m) CRWR_FAIL(CRWR_ERR_INVALID, "scratch too small: need %lld bytes", (long long)(8 * m));
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) CRWR_FAIL(CRWR_ERR_CUDA, "no CUDA device: libcrwr_b200 has no CPU path");
    cudaStream_t st = (cudaStream_t)stream;
    const int tb = 256;
    if (d_scratch) {
        // work proportional to the stored entries of the rows: inverse map of the bins, then one warp per contig
        int32_t* bin_of = (int32_t*)d_scratch;
        int32_t* pos_of = bin_of + m;
        CUDA_TRY(cudaMemsetAsync(bin_of, 0xff, 4 * (size_t)m, st));
        if (n_members > 0) {
            bin_inverse_kernel<<<(int)((n_members + tb - 1) / tb), tb, 0, st>>>(d_bin_ptr, d_bin_members, n_bins, bin_of, pos_of);
            LAUNCH_CHECK();
        }
        const long long grid = (n_rows + kBinRowWarps - 1) / kBinRowWarps;
        if (grid > 2147483647ll) CRWR_FAIL(CRWR_ERR_INVALID, "too many contigs for one launch");
        if (dtype == CRWR_F64)
            bin_weights_rows_kernel<double><<<(int)grid, kBinRowWarps * 32, 0, st>>>(d_indptr, d_indices, (const double*)d_data, d_rows,
                                                                                   n_rows, d_bin_ptr, d_bin_members, n_bins, bin_of,
                                                                                   pos_of, (double*)d_out);
        else
            bin_weights_rows_kernel<float><<<(int)grid, kBinRowWarps * 32, 0, st>>>(d_indptr, d_indices, (const float*)d_data, d_rows,
                                                                                  n_rows, d_bin_ptr, d_bin_members, n_bins, bin_of,
                                                                                  pos_of, (float*)d_out);
        LAUNCH_CHECK();
        return CRWR_OK;
    }
    const long long total = (long long)n_rows * n_bins;
    const long long grid = (total + tb - 1) / tb;
    if (grid > 2147483647ll) CRWR_FAIL(CRWR_ERR_INVALID, "too many (contig, bin) pairs for one launch");
    if (dtype == CRWR_F64)
        bin_weights_kernel<double><<<(int)grid, tb, 0, st>>>(d_indptr, d_indices, (const double*)d_data, d_rows, n_rows, d_bin_ptr,
                                                             d_bin_members, n_bins, (double*)d_out);
    else
        bin_weights_kernel<float><<<(int)grid, tb, 0, st>>>(d_indptr, d_indices, (const float*)d_data, d_rows, n_rows, d_bin_ptr,
                                                            d_bin_members, n_bins, (float*)d_out);
    LAUNCH_CHECK();
    return CRWR_OK;
}

int32_t crwr_profile(crwr_handle h, int32_t enable) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    h->profile = enable != 0;
    return CRWR_OK;
}

int32_t crwr_profile_read(crwr_handle h, double* h_total_ms, int64_t* h_launches, void* stream) {
    if (!h || !h_total_ms || !h_launches) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    double total = 0.0;
    long long n = 0;
    h->pair_ms.clear();
    for (size_t i = 0; i + 1 < h->ev.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess) { total += ms; ++n; h->pair_ms.push_back((double)ms); }
        cudaEventDestroy(h->ev[i]);
        cudaEventDestroy(h->ev[i + 1]);
    }
    h->ev.clear();
    for (size_t i = 0; i < h->tl.size(); ++i) {
        float ms = 0.f;
        if (i > 0 && strcmp(h->tl[i].second, "step start") != 0 &&
            cudaEventElapsedTime(&ms, h->tl[i - 1].first, h->tl[i].first) == cudaSuccess)
            fprintf(stderr, "  %-20s %8.1f us\n", h->tl[i].second, ms * 1e3);
        else if (strcmp(h->tl[i].second, "step start") == 0) fprintf(stderr, "step\n");
    }
    for (auto& e : h->tl) cudaEventDestroy(e.first);
    h->tl.clear();
    *h_total_ms = total;
    *h_launches = n;
    // step phases: [1] propagate + large-row kernel (+ scans), [2] level-1 histogram pass, [3] gather + finish
    double phase[4] = {0, 0, 0, 0};
    for (size_t i = 1; i < h->marks.size(); ++i) {
        float ms = 0.f;
        const int id = h->mark_ids[i];
        if (id > 0 && h->mark_ids[i - 1] == id - 1 && cudaEventElapsedTime(&ms, h->marks[i - 1], h->marks[i]) == cudaSuccess)
            phase[id] += ms;
    }
    for (cudaEvent_t e : h->marks) cudaEventDestroy(e);
    h->marks.clear();
    h->mark_ids.clear();
    h->phase_ms[1] = phase[1]; h->phase_ms[2] = phase[2]; h->phase_ms[3] = phase[3];
    return CRWR_OK;
}

int32_t crwr_profile_steps(crwr_handle h, double* h_ms, int32_t cap) {
    if (!h || !h_ms || cap < 0) CRWR_FAIL(CRWR_ERR_INVALID, "bad argument");
    const int32_t n = (int32_t)h->pair_ms.size() < cap ? (int32_t)h->pair_ms.size() : cap;
    for (int32_t i = 0; i < n; ++i) h_ms[i] = h->pair_ms[i];
    return n;
}

int32_t crwr_profile_phases(crwr_handle h, double* h_ms3) {
    if (!h || !h_ms3) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    h_ms3[0] = h->phase_ms[1]; h_ms3[1] = h->phase_ms[2]; h_ms3[2] = h->phase_ms[3];
    return CRWR_OK;
}

int32_t crwr_set_row_hints(crwr_handle h, int64_t max_row_len, int64_t max_kept) {
    if (!h || max_row_len < 0 || max_kept < 0) CRWR_FAIL(CRWR_ERR_INVALID, "bad row hints");
    if (max_row_len == h->hint_len && max_kept == h->hint_kept && !h->fast_retune) return CRWR_OK;
    h->fast_retune = false;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int e = h->cfg.dtype == CRWR_F64 ? configure_shape<double>(h, max_row_len, max_kept)
                                     : configure_shape<float>(h, max_row_len, max_kept);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "kernel sizing failed: %s", cudaGetErrorString((cudaError_t)e));
    e = h->cfg.dtype == CRWR_F64 ? configure_group_shape<double>(h, max_row_len, max_kept)
                                 : configure_group_shape<float>(h, max_row_len, max_kept);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "row-group kernel sizing failed: %s", cudaGetErrorString((cudaError_t)e));
    e = h->cfg.dtype == CRWR_F64 ? configure_struct_shape<double>(h, max_row_len, max_kept)
                                 : configure_struct_shape<float>(h, max_row_len, max_kept);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "symbolic kernel sizing failed: %s", cudaGetErrorString((cudaError_t)e));
    h->hint_len = max_row_len;
    h->hint_kept = max_kept;
    h->hint_len_eff = max_row_len;
    return CRWR_OK;
}

int32_t crwr_hold_window_select(crwr_handle h, int32_t until_step) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    h->win_hold_until = until_step > 0 ? until_step : 0;
    return CRWR_OK;
}

int32_t crwr_get_shape(crwr_handle h, int64_t* h_out8) {
    if (!h || !h_out8) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    h_out8[0] = h->shape.H; h_out8[1] = h->shape.limit; h_out8[2] = h->shape.kmax; h_out8[3] = h->shape.scap;
    h_out8[4] = h->shape.warps; h_out8[5] = (int64_t)h->shape.per_warp; h_out8[6] = h->grid; h_out8[7] = h->sm_count;
    return CRWR_OK;
}

int32_t crwr_get_group_shape(crwr_handle h, int64_t* h_out8) {
    if (!h || !h_out8) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    const GroupShape& g = h->gshape;
    h_out8[0] = g.H; h_out8[1] = g.limit; h_out8[2] = g.kmax; h_out8[3] = g.ov_cap;
    h_out8[4] = g.G; h_out8[5] = g.warps; h_out8[6] = (int64_t)g.per_row; h_out8[7] = h->prop_kind == 1 ? h->ggrid : 0;
    return CRWR_OK;
}

int32_t crwr_walk_enqueue(crwr_handle h, int32_t n_steps, void* stream) {
    if (!h || !h->walking) CRWR_FAIL(CRWR_ERR_STATE, "crwr_walk_init first");
    if (h->cfg.n_shards != 1)
        CRWR_FAIL(CRWR_ERR_STATE, "crwr_walk_enqueue is for unsharded walks; use the crwr_step_* calls");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (int i = 0; i < n_steps; ++i) {
        int e = h->cfg.dtype == CRWR_F64 ? enqueue_step<double>(h, st) : enqueue_step<float>(h, st);
        if (e) CRWR_FAIL(CRWR_ERR_CUDA, "step launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    return CRWR_OK;
}

// ---- sharded walk: the caller interleaves these with its collectives ----------
#define STEP_PROLOGUE()                                                                  \
    if (!h || !h->walking) CRWR_FAIL(CRWR_ERR_STATE, "crwr_walk_init first");             \
    cudaStream_t st = (cudaStream_t)stream;                                               \
    CUDA_TRY(cudaSetDevice(h->cfg.device));                                               \
    const bool f64 = h->cfg.dtype == CRWR_F64;                                            \
    const int parity = h->steps_enqueued & 1;                                             \
    (void)parity; (void)f64;

// A sharded step takes the window select (one cross-GPU synchronisation, crwr_step_peer_tail) once the walk has settled.
// The candidates of every rank must fit its push regions, so the window has to be narrow: from walk step
// win_step_sharded on (the cut has nearly settled by then; CRWR_WIN_STEP_SHARDED), or earlier when a poll has seen a
// narrow window already.  The state is replicated, so every rank switches in the same step; a window that turns out too
// wide is a miss like any other (the step is repeated the three-barrier way).
static bool sharded_window_step(const crwr_handle_s* h) {
    return h->peer_attached && h->win_step > 0 && h->steps_enqueued >= h->win_step && !h->redo_hist &&
           (h->sharded_window_ok || h->steps_enqueued >= h->win_step_sharded);
}
int32_t crwr_step_window(crwr_handle h) { return (h && h->walking && sharded_window_step(h)) ? 1 : 0; }

int32_t crwr_step_propagate(crwr_handle h, void* stream) {
    STEP_PROLOGUE();
    // steps >= 1: the level-0 histogram is filled by the propagate epilogue (crwr_step_hist(0) then has nothing to do)
    h->hist0_fused = h->steps_enqueued >= 1;
    const bool win = sharded_window_step(h);
    int e = f64 ? launch_propagate<double>(h, parity, false, win, st) : launch_propagate<float>(h, parity, false, win, st);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "propagate launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (!h->peer_attached) {        // with the peer exchange the counters are published by crwr_step_peer_scan / _tail
        pack_exchange_kernel<<<1, 32, 0, st>>>(h->at<WalkState>(h->L.state), xbuf_ptr(h));
        LAUNCH_CHECK();
    }
    return CRWR_OK;
}
int32_t crwr_step_hist(crwr_handle h, int32_t level, void* stream) {
    STEP_PROLOGUE();
    if (level < 0 || level >= kSelectLevels) CRWR_FAIL(CRWR_ERR_INVALID, "level out of range");
    if (level == 0 && h->hist0_fused) return CRWR_OK;        // filled by the propagate epilogue
    int e = f64 ? launch_hist<double>(h, parity ^ 1, level, 0, st) : launch_hist<float>(h, parity ^ 1, level, 0, st);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "hist launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    tl_mark(h, st, "hist pass");
    return CRWR_OK;
}
int32_t crwr_step_scan(crwr_handle h, int32_t level, void* stream) {
    STEP_PROLOGUE();
    if (level < 0 || level >= kSelectLevels) CRWR_FAIL(CRWR_ERR_INVALID, "level out of range");
    int e = f64 ? launch_scan<double>(h, level, st) : launch_scan<float>(h, level, st);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}
int32_t crwr_step_gather(crwr_handle h, void* stream) {
    STEP_PROLOGUE();
    int e = f64 ? launch_gather<double>(h, parity ^ 1, 0, st) : launch_gather<float>(h, parity ^ 1, 0, st);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "gather launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    tl_mark(h, st, "gather");
    return CRWR_OK;
}
int32_t crwr_step_finish(crwr_handle h, const void* d_gathered_blocks, void* stream) {
    STEP_PROLOGUE();
    const int do_select = h->steps_enqueued >= 1 ? 1 : 0;
    if (do_select && h->cfg.n_shards > 1 && !d_gathered_blocks)
        CRWR_FAIL(CRWR_ERR_INVALID, "a sharded step needs the all-gathered candidate blocks");
    const void* g = (do_select && h->cfg.n_shards > 1) ? d_gathered_blocks : nullptr;
    int e = f64 ? launch_finish<double>(h, do_select, g, st) : launch_finish<float>(h, do_select, g, st);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "finish launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    h->steps_enqueued += 1;
    return CRWR_OK;
}
// ---- NVLink peer exchange (sharded walk without NCCL in the step loop) ------------------------------------
int64_t crwr_peer_block_bytes(void) { return 8 * (int64_t)kPeerWords; }

int32_t crwr_peer_attach(crwr_handle h, void* d_block_ptrs, int32_t rank, uint64_t epoch_base) {
    if (!h || !d_block_ptrs) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    if (h->cfg.n_shards < 2 || h->cfg.n_shards > kPeerFlags) CRWR_FAIL(CRWR_ERR_INVALID, "peer exchange needs 2..%d shards", kPeerFlags);
    if (rank < 0 || rank >= h->cfg.n_shards) CRWR_FAIL(CRWR_ERR_INVALID, "rank out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // (the own block's address is read back once per pointer array: the copy synchronises the device)
    unsigned long long* mine = h->peer_block;
    if (!(h->peer_attached && h->peer.blocks == (unsigned long long* const*)d_block_ptrs && h->peer.rank == rank))
        CUDA_TRY(cudaMemcpy(&mine, (unsigned long long**)d_block_ptrs + rank, sizeof(mine), cudaMemcpyDeviceToHost));
    h->peer.blocks = (unsigned long long* const*)d_block_ptrs;
    h->peer.world = h->cfg.n_shards;
    h->peer.rank = rank;
    h->peer_block = mine;
    h->peer_epoch_base = epoch_base;
    h->peer_attached = true;
    return CRWR_OK;
}

// Cross-GPU barrier + sum of the peers' counters / histogram bins + scan.  level -1: counters only (step 0).
int32_t crwr_step_peer_scan(crwr_handle h, int32_t level, void* stream) {
    STEP_PROLOGUE();
    if (!h->peer_attached) CRWR_FAIL(CRWR_ERR_STATE, "crwr_peer_attach first");
    if (level < -1 || level >= kSelectLevels) CRWR_FAIL(CRWR_ERR_INVALID, "level out of range");
    const unsigned long long epoch = h->peer_epoch_base + 3ull * (unsigned long long)h->steps_enqueued + (level == 1 ? 2ull : 1ull);
    unsigned long long* local = h->at<unsigned long long>(h->L.xbuf);
    if (level == 0 && h->scan_push) {
        // push form (flags of its own, one epoch per step)
        const unsigned long long ep = h->peer_epoch_base + (unsigned long long)h->steps_enqueued + 1ull;
        if (f64) launch_k(use_pdl(h), peer_scan_push_kernel<double>, 1, 1024, 0, st, h->at<WalkState>(h->L.state), h->peer, local, h->steps_enqueued & 1, ep);
        else launch_k(use_pdl(h), peer_scan_push_kernel<float>, 1, 1024, 0, st, h->at<WalkState>(h->L.state), h->peer, local, h->steps_enqueued & 1, ep);
    } else if (f64) launch_k(use_pdl(h), peer_scan_kernel<double>, 1, 1024, 0, st, h->at<WalkState>(h->L.state), h->peer, local, level, epoch);
    else launch_k(use_pdl(h), peer_scan_kernel<float>, 1, 1024, 0, st, h->at<WalkState>(h->L.state), h->peer, local, level, epoch);
    LAUNCH_CHECK();
    tl_mark(h, st, level == 1 ? "peer scan 1" : "peer scan 0");
    return CRWR_OK;
}

int32_t crwr_step_peer_finish(crwr_handle h, void* stream) {
    STEP_PROLOGUE();
    if (!h->peer_attached) CRWR_FAIL(CRWR_ERR_STATE, "crwr_peer_attach first");
    const int step = h->steps_enqueued;
    const int do_select = step >= 1 ? 1 : 0;
    const unsigned long long epoch = h->peer_epoch_base + 3ull * (unsigned long long)step + 3ull;
    FinishArgs a = finish_args(h, do_select, nullptr);
    a.xch = h->at<unsigned long long>(h->L.xbuf);         // local buffer: holds the sums peer_scan stored
    unsigned long long* gathered = h->at<unsigned long long>(h->L.gathered);
    if (f64) launch_k(use_pdl(h), peer_finish_kernel<double>, 1, 1024, 0, st, a, h->peer, gathered, step & 1, epoch);
    else launch_k(use_pdl(h), peer_finish_kernel<float>, 1, 1024, 0, st, a, h->peer, gathered, step & 1, epoch);
    LAUNCH_CHECK();
    tl_mark(h, st, "peer finish");
    h->steps_enqueued += 1;
    h->redo_hist = false;
    return CRWR_OK;
}

// Window-select step of a sharded walk: push + one flag barrier + select + finish (peer.cuh).
int32_t crwr_step_peer_tail(crwr_handle h, void* stream) {
    STEP_PROLOGUE();
    if (!h->peer_attached) CRWR_FAIL(CRWR_ERR_STATE, "crwr_peer_attach first");
    if (!sharded_window_step(h)) CRWR_FAIL(CRWR_ERR_STATE, "not a window step (crwr_step_window)");
    const int step = h->steps_enqueued;
    // the push exchange has flags of its own; one epoch per step
    const unsigned long long epoch = h->peer_epoch_base + (unsigned long long)step + 1ull;
    FinishArgs a = finish_args(h, 1, nullptr);
    unsigned long long* local = h->at<unsigned long long>(h->L.xbuf);
    if (f64) launch_k(use_pdl(h), peer_tail_kernel<double>, 1, 1024, 0, st, a, h->peer, local, cand_ptr(h, step), cand_ptr(h, step + 1), step & 1, epoch);
    else launch_k(use_pdl(h), peer_tail_kernel<float>, 1, 1024, 0, st, a, h->peer, local, cand_ptr(h, step), cand_ptr(h, step + 1), step & 1, epoch);
    LAUNCH_CHECK();
    tl_mark(h, st, "peer tail");
    h->steps_enqueued += 1;
    return CRWR_OK;
}

// After a poll that reported a window miss (crwr_status.capacity_overflow bit 2): the walk goes on, the missed step is
// queued again with the histogram select (crwr_step_window answers 0 for it).
int32_t crwr_walk_redo(crwr_handle h, void* stream) {
    if (!h || !h->walking) CRWR_FAIL(CRWR_ERR_STATE, "crwr_walk_init first");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    peer_redo_kernel<<<1, 32, 0, st>>>(h->at<WalkState>(h->L.state));
    LAUNCH_CHECK();
    h->redo_hist = true;
    return CRWR_OK;
}

// n_steps whole steps of a peer-attached shard in one call (what a host loop over the crwr_step_* calls does).
int32_t crwr_walk_enqueue_peer(crwr_handle h, int32_t n_steps, void* stream) {
    if (!h || !h->walking) CRWR_FAIL(CRWR_ERR_STATE, "crwr_walk_init first");
    if (!h->peer_attached) CRWR_FAIL(CRWR_ERR_STATE, "crwr_peer_attach first");
    for (int i = 0; i < n_steps; ++i) {
        int32_t rc = CRWR_OK;
        if (sharded_window_step(h)) {
            if ((rc = crwr_step_propagate(h, stream)) != CRWR_OK) return rc;
            if ((rc = crwr_step_peer_tail(h, stream)) != CRWR_OK) return rc;
            continue;
        }
        const bool first = h->steps_enqueued == 0;
        if ((rc = crwr_step_propagate(h, stream)) != CRWR_OK) return rc;
        if (first) {
            if ((rc = crwr_step_peer_scan(h, -1, stream)) != CRWR_OK) return rc;
        } else {
            if ((rc = crwr_step_peer_scan(h, 0, stream)) != CRWR_OK) return rc;
            if ((rc = crwr_step_hist(h, 1, stream)) != CRWR_OK) return rc;
            if ((rc = crwr_step_peer_scan(h, 1, stream)) != CRWR_OK) return rc;
            if ((rc = crwr_step_gather(h, stream)) != CRWR_OK) return rc;
        }
        if ((rc = crwr_step_peer_finish(h, stream)) != CRWR_OK) return rc;
    }
    return CRWR_OK;
}

void* crwr_exchange_ptr(crwr_handle h, int32_t which) {
    if (!h) return nullptr;
    const Layout& L = h->L;
    if (which == 0) return h->ws + L.xbuf;
    if (which == 1) return h->ws + L.xbuf + 8 * (size_t)(kXExtras + kSelectBins);
    if (which == 2) return h->ws + L.cand;
    return nullptr;
}
int64_t crwr_exchange_bytes(crwr_handle h, int32_t which) {
    if (!h) return -1;
    if (which == 0) return 8 * (int64_t)(kXExtras + kSelectBins);
    if (which == 1) return 8 * (int64_t)kSelectBins;
    if (which == 2) return 8 * (int64_t)(kCandHeader + h->L.cand_cap);
    return -1;
}

int32_t crwr_walk_poll(crwr_handle h, crwr_status* out, void* stream) {
    if (!h || !out) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->h_state, h->ws + h->L.state, sizeof(WalkState), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const WalkState& s = *h->h_state;
    out->steps_done = s.step;
    out->stop_reason = (s.done && !s.win_missed) ? s.stop_reason : CRWR_STOP_RUNNING;
    out->capacity_overflow = s.capacity_overflow | (s.select_overflow << 1) | (s.win_missed << 2);
    h->sharded_window_ok = s.win_on != 0 && s.win_narrow != 0;
    out->fallback_rows = s.last_fallback_rows;
    out->nnz_iterate = (int64_t)s.last_nnz_new;
    out->max_row_len = s.last_max_row_len;
    out->max_kept = s.last_max_kept;
    out->last_cut = s.cut;
    out->last_delta = s.last_delta;
    if (s.done && !s.win_missed && s.stop_reason == CRWR_STOP_RUNNING) out->stop_reason = CRWR_STOP_MAX_STEPS;
    // the host-side step counter follows the device: steps enqueued past the stop were skipped
    if (s.done) h->steps_enqueued = s.step;
    if (s.last_fallback_rows > 0) h->fb_seen = true;
    if (s.big_rows > 0) h->big_seen = true;
    {
        // (a sharded walk's count is the sum over all ranks, its rows are the markers)
        const long long rows = h->cfg.n_shards > 1 ? h->cfg.n_markers : h->cfg.row_end - h->cfg.row_begin;
        h->mean_row = (long long)(s.last_nnz_new / (unsigned long long)(rows > 0 ? rows : 1));
    }
    if (s.hot_sum > 0) {
        const long long rows = h->cfg.row_end - h->cfg.row_begin;
        const long long avg = (long long)(s.hot_sum / (unsigned long long)(rows > 0 ? rows : 1));
        if (h->avg_hot == 0 || avg > h->avg_hot + h->avg_hot / 4 || avg < h->avg_hot - h->avg_hot / 4) {
            h->avg_hot = avg;
            h->fast_retune = true;
        }
    }
    if (s.fb_unhandled) {
        h->dense_always = true;
        CRWR_FAIL(CRWR_ERR_RETRY, "rows were deferred in a step without a large-row kernel: walk again (the kernel is now always queued)");
    }
    if (s.debug_error) CRWR_FAIL(CRWR_ERR_STATE, "internal bounds check %d failed (CRWR_DEBUG_CHECKS build)", s.debug_error);
    if (s.peer_timeout) CRWR_FAIL(CRWR_ERR_STATE, "a cross-GPU barrier timed out: a peer rank did not arrive");
    if (s.capacity_overflow) CRWR_FAIL(CRWR_ERR_CAPACITY, "iterate buffer overflow (capacity %lld entries)", (long long)h->cfg.iterate_cap);
    if (s.select_overflow) CRWR_FAIL(CRWR_ERR_CAPACITY, "percentile candidate buffer overflow");
    if (s.order_hazard) CRWR_FAIL(CRWR_ERR_STATE, "a walk step ended without a percentile cut although its rows were stored unordered (set CRWR_PROP=warp)");
    return CRWR_OK;
}

int32_t crwr_walk_trace(crwr_handle h, crwr_step_record* out, int32_t cap, void* stream) {
    if (!h || !out) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    int n = cap < kMaxTrace ? cap : kMaxTrace;
    CUDA_TRY(cudaMemcpyAsync(out, h->ws + h->L.trace, sizeof(crwr_step_record) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CRWR_OK;
}

int32_t crwr_result_count(crwr_handle h, int64_t* h_nnz, void* stream) {
    if (!h || !h_nnz) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int e = h->cfg.dtype == CRWR_F64 ? count_rows<double>(h, (int)h->cfg.n_markers, h_nnz, (cudaStream_t)stream)
                                     : count_rows<float>(h, (int)h->cfg.n_markers, h_nnz, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "result count failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}
int32_t crwr_result_write(crwr_handle h, int64_t* d_indptr, int32_t* d_indices, void* d_data, void* stream) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    int e = h->cfg.dtype == CRWR_F64 ? write_rows<double>(h, (int)h->cfg.n_markers, d_indptr, d_indices, d_data, (cudaStream_t)stream)
                                     : write_rows<float>(h, (int)h->cfg.n_markers, d_indptr, d_indices, d_data, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "result write failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}
int32_t crwr_iterate_count(crwr_handle h, int64_t* h_nnz, void* stream) {
    if (!h || !h_nnz) CRWR_FAIL(CRWR_ERR_INVALID, "null argument");
    int e = h->cfg.dtype == CRWR_F64 ? count_rows<double>(h, 0x7fffffff, h_nnz, (cudaStream_t)stream)
                                     : count_rows<float>(h, 0x7fffffff, h_nnz, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "iterate count failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}
int32_t crwr_iterate_write(crwr_handle h, int64_t* d_indptr, int32_t* d_indices, void* d_data, void* stream) {
    if (!h) CRWR_FAIL(CRWR_ERR_INVALID, "null handle");
    int e = h->cfg.dtype == CRWR_F64 ? write_rows<double>(h, 0x7fffffff, d_indptr, d_indices, d_data, (cudaStream_t)stream)
                                     : write_rows<float>(h, 0x7fffffff, d_indptr, d_indices, d_data, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "iterate write failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}

}  // extern "C"

// ---- symmetrise ------------------------------------------------------------------------------
namespace {
template <typename T>
size_t sym_layout(int64_t m, int64_t nnz, SymScratch<T>* sc, unsigned char* base) {
    size_t off = 0;
    size_t a;
    a = bump(off, 4 * (size_t)(m + 1)); if (sc) sc->t_cnt = (int32_t*)(base + a);
    a = bump(off, 8 * (size_t)(m + 1)); if (sc) sc->t_ptr = (int64_t*)(base + a);
    a = bump(off, 4 * (size_t)(m + 1)); if (sc) sc->t_fill = (int32_t*)(base + a);
    a = bump(off, 4 * (size_t)(nnz + 1)); if (sc) sc->t_col = (int32_t*)(base + a);
    a = bump(off, sizeof(T) * (size_t)(nnz + 1)); if (sc) sc->t_val = (T*)(base + a);
    a = bump(off, 4 * (size_t)(m + 1)); if (sc) sc->e_cnt = (int32_t*)(base + a);
    a = bump(off, 8 * (size_t)(m + 1)); if (sc) sc->e_ptr = (int64_t*)(base + a);
    a = bump(off, sizeof(T) * (size_t)(m + 1)); if (sc) sc->b = (T*)(base + a);
    a = bump(off, 8); if (sc) sc->total = (long long*)(base + a);
    return (off + 255) & ~(size_t)255;
}

template <typename T>
int sym_count(int64_t m, const int64_t* ip, const int32_t* ix, const T* data, int64_t nnz, void* scratch,
              int64_t* h_out, cudaStream_t st) {
    SymScratch<T> sc;
    sym_layout<T>(m, nnz, &sc, (unsigned char*)scratch);
    if (cudaMemsetAsync(sc.t_cnt, 0, 4 * (size_t)(m + 1), st) != cudaSuccess) return -1;
    if (cudaMemsetAsync(sc.t_fill, 0, 4 * (size_t)(m + 1), st) != cudaSuccess) return -1;
    const int blocks = (int)((m * 32 + 255) / 256);
    sym_count_cols_kernel<T><<<blocks, 256, 0, st>>>((int)m, ip, ix, sc);
    ++g_launches;
    exclusive_scan_kernel<int64_t><<<1, 1024, 0, st>>>(sc.t_cnt, sc.t_ptr, m, nullptr);
    ++g_launches;
    sym_scatter_kernel<T><<<blocks, 256, 0, st>>>((int)m, ip, ix, data, sc);
    ++g_launches;
    sym_sort_count_kernel<T><<<blocks, 256, 0, st>>>((int)m, ip, ix, sc);
    ++g_launches;
    exclusive_scan_kernel<int64_t><<<1, 1024, 0, st>>>(sc.e_cnt, sc.e_ptr, m, sc.total);
    ++g_launches;
    long long total = 0;
    if (cudaMemcpyAsync(&total, sc.total, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    if (cudaGetLastError() != cudaSuccess) return -1;
    *h_out = total;
    return 0;
}

template <typename T>
int sym_write(int64_t m, const int64_t* ip, const int32_t* ix, const T* data, int64_t nnz, void* scratch,
              int64_t* out_ip, int32_t* out_ix, T* out_data, cudaStream_t st) {
    SymScratch<T> sc;
    sym_layout<T>(m, nnz, &sc, (unsigned char*)scratch);
    if (cudaMemcpyAsync(out_ip, sc.e_ptr, 8 * (size_t)(m + 1), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
    const int blocks = (int)((m * 32 + 255) / 256);
    sym_merge_kernel<T><<<blocks, 256, 0, st>>>((int)m, ip, ix, data, sc, sc.e_ptr, out_ix, out_data);
    ++g_launches;
    sym_scale_kernel<T><<<blocks, 256, 0, st>>>((int)m, sc.e_ptr, out_ix, out_data, sc.b);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}  // namespace

extern "C" {

int64_t crwr_symmetrise_scratch_bytes(int64_t m, int64_t nnz) {
    if (m < 1 || nnz < 0) return -1;
    return (int64_t)sym_layout<double>(m, nnz, nullptr, nullptr);
}

int32_t crwr_symmetrise_count(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices, const void* d_data,
                              int64_t nnz, void* d_scratch, int64_t scratch_bytes, int64_t* h_nnz_out, void* stream) {
    if (m < 1 || nnz < 0 || !d_scratch || !h_nnz_out) CRWR_FAIL(CRWR_ERR_INVALID, "bad argument");
    if (scratch_bytes < crwr_symmetrise_scratch_bytes(m, nnz)) CRWR_FAIL(CRWR_ERR_INVALID, "scratch too small");
    int e = dtype == CRWR_F64 ? sym_count<double>(m, d_indptr, d_indices, (const double*)d_data, nnz, d_scratch, h_nnz_out, (cudaStream_t)stream)
                              : sym_count<float>(m, d_indptr, d_indices, (const float*)d_data, nnz, d_scratch, h_nnz_out, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "symmetrise count failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}

int32_t crwr_symmetrise_write(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices, const void* d_data,
                              int64_t nnz, void* d_scratch, int64_t scratch_bytes, int64_t* d_out_indptr, int32_t* d_out_indices,
                              void* d_out_data, void* stream) {
    if (m < 1 || nnz < 0 || !d_scratch) CRWR_FAIL(CRWR_ERR_INVALID, "bad argument");
    if (scratch_bytes < crwr_symmetrise_scratch_bytes(m, nnz)) CRWR_FAIL(CRWR_ERR_INVALID, "scratch too small");
    int e = dtype == CRWR_F64 ? sym_write<double>(m, d_indptr, d_indices, (const double*)d_data, nnz, d_scratch, d_out_indptr, d_out_indices, (double*)d_out_data, (cudaStream_t)stream)
                              : sym_write<float>(m, d_indptr, d_indices, (const float*)d_data, nnz, d_scratch, d_out_indptr, d_out_indices, (float*)d_out_data, (cudaStream_t)stream);
    if (e) CRWR_FAIL(CRWR_ERR_CUDA, "symmetrise write failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRWR_OK;
}

}  // extern "C"
