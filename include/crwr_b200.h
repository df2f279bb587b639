/*
 * crwr_b200.h - C ABI of libcrwr_b200.so: constrained random walk with restart (CRWR)
 * imputation on NVIDIA B200 (sm_100a).
 *
 * The reference (dyxstat/ImputeCC, pure Python) has no FFI for this path; its boundary is two
 * Python functions.  The entry points below are what a binding of that boundary needs:
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  ------------------------------------
 *   Script/imputation.py:112-119  permute + strip diag +  crwr_build_transition
 *                                 self-loops + row scale
 *   Script/utility.py:270-279     random_walk_cpu entry,  crwr_set_transition (caller-built P),
 *                                 restart matrix I, Q=I   crwr_walk_init
 *   Script/utility.py:281-304     the walk loop           crwr_walk_enqueue / crwr_walk_poll
 *                                                         (per-phase calls for the sharded walk)
 *   Script/utility.py:306-307     Q[:m,:m]                crwr_result_count / crwr_result_write
 *   Script/imputation.py:122-127  E=Q+Q^T, D^-1/2 E D^-1/2 crwr_symmetrise_count / _write
 *   Script/utility.py:466-480     match_contigs edge       crwr_bin_contact_weights
 *                                 weights (next row, 8f)
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer (e.g. torch.Tensor.data_ptr()); h_* is host.
 *   - the library never allocates device memory: the caller passes one workspace block whose
 *     size crwr_workspace_bytes() reports, plus explicit output buffers (two-call pattern:
 *     *_count then *_write, because output sizes are data dependent).
 *   - stream is a cudaStream_t passed as void*; calls only enqueue work unless stated.
 *   - return value: 0 = CRWR_OK, negative = error; crwr_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef CRWR_B200_H
#define CRWR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRWR_ABI_VERSION 1
#define CRWR_SHARD_CAND 8192   /* percentile candidates one shard may contribute per step */

enum { CRWR_F32 = 0, CRWR_F64 = 1 };

enum {
    CRWR_OK = 0,
    CRWR_ERR_INVALID = -1,     /* bad argument */
    CRWR_ERR_CUDA = -2,        /* CUDA runtime error */
    CRWR_ERR_CAPACITY = -3,    /* iterate buffers too small: recreate with a larger capacity */
    CRWR_ERR_STATE = -4,       /* call out of order */
    CRWR_ERR_RETRY = -5        /* the walk must be run again from crwr_walk_init (the handle has adapted itself) */
};

enum { CRWR_STOP_RUNNING = 0, CRWR_STOP_TOLERANCE = 1, CRWR_STOP_PLATEAU = 2, CRWR_STOP_MAX_STEPS = 3 };

typedef struct crwr_handle_s* crwr_handle;

typedef struct crwr_config {
    int32_t dtype;              /* CRWR_F32 (reference arithmetic, imputation.py:119) or CRWR_F64 */
    int32_t device;             /* CUDA device ordinal */
    int64_t n_contigs;          /* N: order of the contact matrix */
    int64_t n_markers;          /* m: restart rows of the whole job (utility.py:274) */
    int64_t row_begin;          /* this shard walks restart rows [row_begin, row_end) */
    int64_t row_end;
    int64_t transition_nnz_cap; /* capacity for nnz(P) */
    int64_t iterate_cap;        /* capacity, in entries, of EACH of the two iterate buffers */
    int32_t n_shards;           /* ranks sharing the walk (1 = unsharded: rows must be [0, m)) */
    int32_t reserved;
} crwr_config;

/* One record per walk step; mirrors oracle StepTrace. */
typedef struct crwr_step_record {
    int64_t nnz_in;     /* entries of the thresholded iterate entering the step */
    int64_t nnz_new;    /* entries of (1-rp) Q P + rp I */
    double cut;         /* percentile value; NaN when no cut was computed (step 0) */
    double delta;       /* ||Q_prev - Q_new||_F */
    int32_t cut_applied;
    int32_t plateau_hits;
    int32_t fallback_rows;  /* rows that took the large-row path in this step */
    int32_t max_row_len;    /* longest row of the step's product */
    int32_t max_kept;       /* longest kept operand row read by the step */
    int32_t fast_rows;      /* rows of this shard computed through their cached structure (structure.cuh) */
} crwr_step_record;

typedef struct crwr_status {
    int32_t steps_done;
    int32_t stop_reason;        /* CRWR_STOP_* */
    int32_t capacity_overflow;  /* bit 0: an iterate buffer overflowed, bit 1: candidate buffer overflowed (results invalid);
                                 * bit 2: sharded walk only - the select window missed, call crwr_walk_redo */
    int32_t fallback_rows;      /* rows that took the large-row path in the last step */
    int64_t nnz_iterate;        /* entries stored by the last step's product (before its cut) */
    int64_t max_row_len;        /* longest stored row */
    int64_t max_kept;           /* longest kept (post-cut) operand row read by the last step */
    double last_cut;
    double last_delta;
} crwr_status;

int32_t crwr_abi_version(void);
const char* crwr_last_error(void);

/* Bytes of device workspace crwr_create needs for this configuration. */
int64_t crwr_workspace_bytes(const crwr_config* cfg);

int32_t crwr_create(crwr_handle* out, const crwr_config* cfg, void* d_workspace, int64_t workspace_bytes);
int32_t crwr_destroy(crwr_handle h);

/* imputation.py:112-119.  NormCC CSR (int64 indptr, int32 indices, float64 data) in ORIGINAL
 * contig order plus d_new_of_old[N] (position of each contig in the markers-first order of
 * imputation.py:98-110).  Builds P (CSR, markers-first index space, sorted rows) in the
 * workspace.  h_nnz_out (optional, host) receives nnz(P) after an internal stream sync. */
int32_t crwr_build_transition(crwr_handle h, const int64_t* d_indptr, const int32_t* d_indices,
                              const double* d_data, int64_t nnz, const int32_t* d_new_of_old,
                              int64_t* h_nnz_out, void* stream);

/* B-inner (utility.py:270): caller supplies P exactly as random_walk_cpu would receive it
 * (CSR, int32 indptr/indices, values of the handle's dtype, markers in rows/cols 0..m-1). */
int32_t crwr_set_transition(crwr_handle h, const int32_t* d_indptr, const int32_t* d_indices,
                            const void* d_data, int64_t nnz, void* stream);

/* Copy P out (for tests): arrays sized N+1 / nnz / nnz. */
int32_t crwr_get_transition(crwr_handle h, int32_t* d_indptr, int32_t* d_indices, void* d_data, void* stream);

/* utility.py:275-280: Q = I, counters cleared.  perc in [0,100]; tol 0.01 and max_steps 500 in
 * the reference (imputation.py:120, utility.py:281). */
int32_t crwr_walk_init(crwr_handle h, double rp, double perc, double tol, int32_t max_steps, void* stream);

/* Sizing hints for the propagation kernel: the longest output row and the longest kept operand
 * row to expect.  They size the per-row shared-memory accumulator; rows that do not fit take the
 * large-row path, so hints affect speed only, never results.  Host policy derives them from
 * crwr_status.max_row_len / max_kept. */
int32_t crwr_set_row_hints(crwr_handle h, int64_t max_row_len, int64_t max_kept);
/* Host policy for the settled-walk form of the select (one single-CTA kernel per step working on a window around the
 * previous cut): no such step before walk step `until_step` (0: the library's default, from step 6 on).  While the cut
 * still moves by a large factor per step no window can be set and that kernel would have to select over the whole value
 * array by itself; the multi-CTA histogram select of the first steps is kept instead.  Speed only, never results;
 * cleared by crwr_walk_init. */
int32_t crwr_hold_window_select(crwr_handle h, int32_t until_step);
int64_t crwr_transition_nnz(crwr_handle h);
/* Current kernel sizing (diagnostics): {H, limit, kmax, scap, warps/CTA, smem bytes/warp, grid, SMs}. */
int32_t crwr_get_shape(crwr_handle h, int64_t* h_out8);
/* Sizing of the row-group propagate kernel (steps >= 1; diagnostics):
 * {table slots/row, max columns/row, max kept/row, overflow entries/row, lanes/row, warps/CTA, smem bytes/row, grid}. */
int32_t crwr_get_group_shape(crwr_handle h, int64_t* h_out8);

/* Enqueue up to n_steps walk steps (utility.py:282-304).  Stop rules are evaluated on the
 * device; steps enqueued after the walk has stopped are no-ops.  Single-shard walks only. */
int32_t crwr_walk_enqueue(crwr_handle h, int32_t n_steps, void* stream);

/* Sharded walk (one process per GPU, rows [row_begin,row_end) each).  One step is
 *   propagate; hist(0); ALLREDUCE-SUM-u64 exchange 0; scan(0); hist(1); ALLREDUCE exchange 1;
 *   scan(1); gather; ALLGATHER exchange 2 (candidate block); finish(gathered blocks)
 * Step 0 has no cut (utility.py:286): propagate; ALLREDUCE exchange 0; finish(NULL).
 * Exchange 0 = [8 counters incl. the fixed-point delta^2 | level-0 histogram], exchange 1 =
 * level-1 histogram, exchange 2 = [count, successor key, CRWR_SHARD_CAND keys] (all uint64).
 * Every rank then derives the identical cut and stop decision.  See INTEGRATION.md. */
int32_t crwr_step_propagate(crwr_handle h, void* stream);
int32_t crwr_step_hist(crwr_handle h, int32_t level, void* stream);
int32_t crwr_step_scan(crwr_handle h, int32_t level, void* stream);
int32_t crwr_step_gather(crwr_handle h, void* stream);
int32_t crwr_step_finish(crwr_handle h, const void* d_gathered_blocks, void* stream);
void* crwr_exchange_ptr(crwr_handle h, int32_t which);     /* device pointer inside the workspace */
int64_t crwr_exchange_bytes(crwr_handle h, int32_t which);

/* Sharded walk over NVLink peer memory (no collective library in the step loop).  Every rank allocates one
 * block of crwr_peer_block_bytes() in memory that all GPUs of the group have mapped (torch symmetric memory /
 * CUDA IPC), zeroes it once, and attaches the device array of the n_shards block pointers (rank order).
 * epoch_base must be the same on every rank and larger for every new walk on the same blocks (e.g. walk
 * number << 32).  A step is then
 *   crwr_step_propagate; crwr_step_peer_scan(0); crwr_step_hist(1); crwr_step_peer_scan(1);
 *   crwr_step_gather; crwr_step_peer_finish                 (step 0: propagate; peer_scan(-1); peer_finish)
 * peer_scan = cross-GPU flag barrier + sum of the peers' counters / histogram bins read over NVLink + scan;
 * peer_finish = barrier + copy of the peers' candidate blocks + finish.  Results are bit-identical to the
 * collective-based sequence and to the unsharded walk. */
int64_t crwr_peer_block_bytes(void);
int32_t crwr_peer_attach(crwr_handle h, void* d_block_ptrs, int32_t rank, uint64_t epoch_base);
int32_t crwr_step_peer_scan(crwr_handle h, int32_t level, void* stream);
int32_t crwr_step_peer_finish(crwr_handle h, void* stream);
/* Once the walk has settled (the cut moves by a few percent per step) a sharded step needs ONE cross-GPU
 * synchronisation: the propagation classifies its values against a window around the expected cut (replicated state),
 * and crwr_step_peer_tail pushes [step counters | values below the window | the keys inside it] into every peer's
 * block, raises one flag per peer, waits for theirs, and derives the cut and the stop decision from local memory:
 *   if (crwr_step_window(h)) { crwr_step_propagate; crwr_step_peer_tail; } else { ...the sequence above... }
 * A window that misses is no error: every rank rewinds the step and stops; crwr_walk_poll then reports bit 2 of
 * crwr_status.capacity_overflow, and after crwr_walk_redo the missed step is queued again the three-barrier way
 * (crwr_step_window answers 0 for it).  Replaces the exchange part of utility.py:283-304 on several GPUs. */
int32_t crwr_step_window(crwr_handle h);
int32_t crwr_step_peer_tail(crwr_handle h, void* stream);
int32_t crwr_walk_redo(crwr_handle h, void* stream);
/* n_steps whole steps of a peer-attached shard (either form, as crwr_step_window decides) in one call. */
int32_t crwr_walk_enqueue_peer(crwr_handle h, int32_t n_steps, void* stream);

/* Synchronises the stream and fills *h_out. */
int32_t crwr_walk_poll(crwr_handle h, crwr_status* h_out, void* stream);
/* Copies the per-step records (steps_done of them, at most cap) to host; syncs. */
int32_t crwr_walk_trace(crwr_handle h, crwr_step_record* h_out, int32_t cap, void* stream);

/* utility.py:306-307: the columns < m of the final (cut) iterate, rows of this shard.
 * count: syncs and returns the number of entries in *h_nnz.  write: CSR with int64 indptr
 * (rows_local+1), int32 indices, values of the handle's dtype; columns ascending per row. */
int32_t crwr_result_count(crwr_handle h, int64_t* h_nnz, void* stream);
int32_t crwr_result_write(crwr_handle h, int64_t* d_indptr, int32_t* d_indices, void* d_data, void* stream);
/* The whole m_local x N iterate after the last cut (for tests / checkpoints). */
int32_t crwr_iterate_count(crwr_handle h, int64_t* h_nnz, void* stream);
int32_t crwr_iterate_write(crwr_handle h, int64_t* d_indptr, int32_t* d_indices, void* d_data, void* stream);

/* imputation.py:122-127 on an m x m CSR matrix Q (sorted rows): E = Q + Q^T, d = colsum(E),
 * d[d==0] = 1, E = D^-1/2 E D^-1/2.  Stateless apart from the scratch block the caller passes
 * (crwr_symmetrise_scratch_bytes).  count syncs; write emits CSR with sorted rows. */
int64_t crwr_symmetrise_scratch_bytes(int64_t m, int64_t nnz);
int32_t crwr_symmetrise_count(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices,
                              const void* d_data, int64_t nnz, void* d_scratch, int64_t scratch_bytes,
                              int64_t* h_nnz_out, void* stream);
int32_t crwr_symmetrise_write(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices,
                              const void* d_data, int64_t nnz, void* d_scratch, int64_t scratch_bytes,
                              int64_t* d_out_indptr, int32_t* d_out_indices, void* d_out_data, void* stream);

/* utility.py:466-480 (match_contigs, edge weights of one iteration): for every contig to bin
 * d_rows[r] and every bin b (members d_bin_members[d_bin_ptr[b] .. d_bin_ptr[b+1]) in the order of the reference's
 * bin list, never empty, bins disjoint), out[r * n_bins + b] = -(sum over the members j of E[row, j]) / size(b), summed
 * in that order in the matrix dtype.  E: m x m CSR with sorted rows (the output of crwr_symmetrise_write), local contig
 * indices throughout; n_members = d_bin_ptr[n_bins].  d_scratch: crwr_bin_contact_scratch_bytes(m) bytes for the
 * inverse map of the bins (work then proportional to the stored entries of the rows), or NULL (one thread per
 * (contig, bin) pair with binary searches; bins may then overlap).  Stateless; only enqueues kernels. */
int64_t crwr_bin_contact_scratch_bytes(int64_t m);
int32_t crwr_bin_contact_weights(int32_t dtype, int64_t m, const int64_t* d_indptr, const int32_t* d_indices, const void* d_data,
                                 int64_t n_rows, const int32_t* d_rows, int64_t n_bins, const int64_t* d_bin_ptr,
                                 const int32_t* d_bin_members, int64_t n_members, void* d_scratch, int64_t scratch_bytes,
                                 void* d_out, void* stream);

/* Measurement aid: with profiling enabled every launch of the propagation kernel (the dominant
 * kernel) is bracketed by CUDA events on the launching stream; read returns the summed duration of
 * the launches since the last read (launches skipped after the stop rule count as ~0) and clears. */
int32_t crwr_profile(crwr_handle h, int32_t enable);
int32_t crwr_profile_read(crwr_handle h, double* h_total_ms, int64_t* h_launches, void* stream);
/* After crwr_profile_read: summed durations (ms) of the step phases since the previous read:
 * [0] propagate + large-row kernel, [1] level-1 histogram pass, [2] candidate gather + finish. */
/* Per walk step of the last profiled walk: milliseconds of its propagation launches (what crwr_profile_read sums up).
 * Returns the number of entries written (at most cap). */
int32_t crwr_profile_steps(crwr_handle h, double* h_ms, int32_t cap);
int32_t crwr_profile_phases(crwr_handle h, double* h_ms3);

/* Launch counter: kernels this library launched since the last reset (bench gpu_launches). */
int64_t crwr_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* CRWR_B200_H */
