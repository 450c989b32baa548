/*
 * viloca_b200.h — C ABI of the B200-native CAVI engine for VILOCA's local haplotype inference.
 *
 * This is the drop-in boundary for ONE path of cbg-ethz/VILOCA: the coordinate-ascent
 * variational inference loop of
 *   viloca/local_haplotype_inference/use_quality_scores/cavi.py:70-210   (run_cavi)
 *   viloca/local_haplotype_inference/learn_error_params/cavi.py:71-255   (run_cavi)
 * together with the update equations (update_eqs.py), the ELBO (elbo_eqs.py) and the
 * finishing contraction of analyze_results.py:74-80.  The reference has no FFI for this
 * path (it is NumPy); the seam this library replaces is the Python call
 *   cavi.run_cavi(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary,
 *                 reads_weights, reads_log_error_proba, start_id, output_dir,
 *                 convergence_threshold, record_history, rng, seed_number)
 * (use_quality_scores/cavi.py:70-84) batched over all windows of a genome and all random
 * restarts.  Conventions follow the reference's existing native module libshorah
 * (lib/src/bindings.cpp:13-34): plain scalars and pointers in, int return, 0 = success.
 * Unlike libshorah there is NO process-global state: every call takes a batch handle and
 * handles are independent (one per GPU / host thread).
 *
 * Soft CAVI exits ("ELBO converged.", "Error: ELBO is decreasing.", "Error: ELBO is nan.",
 * use_quality_scores/cavi.py:155-165) are per-restart status codes, never error returns.
 * A non-zero return means a hard error (bad argument, CUDA failure); vl_last_error() gives
 * the message for the calling thread.
 *
 * Memory: the caller owns every buffer.  Host pointers are read/written synchronously with
 * respect to the call that takes them.  The device workspace is allocated by the caller
 * (PyTorch owns it in the Python host layer) and must outlive the batch.
 */
#ifndef VILOCA_B200_H
#define VILOCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_ABI_VERSION 2
#define VL_N_BASES 5          /* alphabet "ACGT-" (shotgun.py:197, run_dpm_mfa.py:26) */
#define VL_CODE_N 255         /* read / reference character outside the alphabet      */
#define VL_MAX_CLUSTERS 128   /* K = --n_max_haplotypes (cli.py:179-181), default 100  */

/* inference mode: which reference package the numbers must match */
#define VL_MODE_UQS 0         /* use_quality_scores  */
#define VL_MODE_LEP 1         /* learn_error_params  */

/* per-restart exit state (exit_message of run_cavi) */
#define VL_STATUS_RUNNING    0
#define VL_STATUS_CONVERGED  1  /* "ELBO converged."            cavi.py:163-165 */
#define VL_STATUS_DECREASING 2  /* "Error: ELBO is decreasing." cavi.py:160-162 */
#define VL_STATUS_NAN        3  /* "Error: ELBO is nan."        cavi.py:155-159 */
#define VL_STATUS_MAXITER    4  /* iteration cap hit (the reference has no cap) */

/* layout of the per-restart scalar blocks (doubles) */
#define VL_INIT_STRIDE 8      /* a0, b0, mean_log_gamma[0], mean_log_gamma[1], c0, d0, mean_log_theta[0], mean_log_theta[1] */
#define VL_SCALAR_STRIDE 16   /* see vl_scalar_index */
enum vl_scalar_index {
  VL_S_ELBO = 0, VL_S_GAMMA_A, VL_S_GAMMA_B, VL_S_MLG0, VL_S_MLG1,
  VL_S_THETA_C, VL_S_THETA_D, VL_S_MLT0, VL_S_MLT1,
  VL_S_DSUM_ALPHA, VL_S_DSUM_AB, VL_S_DSUM_CD,
  VL_S_TERM_DATA, VL_S_TERM_PI, VL_S_TERM_CLUSTER, VL_S_TERM_HAPLO
};

/*
 * One window = one call of run_dpm_mfa.main (use_quality_scores/run_dpm_mfa.py:18-31).
 * Row order of the reads is the first-seen order of preparation.py:56-76 — it defines
 * which row of init_cluster belongs to which read.
 */
typedef struct vl_window_desc {
  int32_t mode;          /* VL_MODE_*                                                    */
  int32_t n_reads;       /* N: unique reads (rows)                                       */
  int32_t length;        /* L: window length                                             */
  int32_t n_clusters;    /* K                                                            */
  int32_t n_starts;      /* R: random restarts (cavi.py:26-67)                           */
  int32_t max_iter;      /* iteration cap per restart, > 0                               */
  double alpha0;         /* Dirichlet concentration (cli.py:112-114)                     */
  double conv_thres;     /* |dELBO| threshold, uqs only (cli.py:183-185); lep uses 1e-3  */
  const uint8_t* codes;        /* [N*L] 0..4 = "ACGT-", VL_CODE_N otherwise (one-hot row of zeros) */
  const double* log_hit;       /* uqs: [N*L] log theta_nl               (preparation.py:123-157)   */
  const double* log_off;       /* uqs: [N*L] log((1-theta_nl)/(B-1)); both NULL for lep            */
  const double* weights;       /* [N] read multiplicities (reads_weights)                          */
  const uint8_t* ref_codes;    /* [L] reference, same coding                                       */
  const double* init_cluster;  /* [R*N*K] initial mean_cluster (initialization.py:46-54)           */
  const double* init_scalars;  /* [R*VL_INIT_STRIDE]                                               */
} vl_window_desc;

/* Caller-allocated result buffers of one window; any pointer may be NULL to skip it. */
typedef struct vl_window_result {
  double* elbo;            /* [R] final ELBO                                             */
  int32_t* n_iterations;   /* [R] `iter` as reported in dict_result["n_iterations"]      */
  int32_t* n_updates;      /* [R] number of update() calls actually performed            */
  int32_t* status;         /* [R] VL_STATUS_*                                            */
  int32_t* converged;      /* [R] the sticky `converged` flag                            */
  int32_t* history_len;    /* [R]                                                        */
  double* history;         /* [R*history_stride] history_elbo                            */
  int32_t history_stride;
  int32_t state_restart;   /* in: restart whose state to copy, or -1 = best ELBO         */
  int32_t best_restart;    /* out: argmax ELBO, first on ties (run_dpm_mfa.py:83-95)     */
  double* mean_haplo;      /* [K*L*B]  mean_haplo of state_restart, C order (K,L,B)      */
  double* mean_cluster;    /* [N*K]                                                      */
  double* alpha;           /* [K]                                                        */
  double* mean_log_pi;     /* [K]                                                        */
  double* scalars;         /* [VL_SCALAR_STRIDE]                                         */
} vl_window_result;

typedef struct vl_batch vl_batch;

int vl_abi_version(void);
const char* vl_last_error(void);

/* Number of CUDA devices visible, or a negative value if the runtime is unusable. */
int vl_device_count(void);

/*
 * Device bytes a batch over these windows needs.  Of the host pointers in the descs only
 * `codes` is read: the number of entries whose base differs from the most frequent base of
 * their position sizes the batch's minority lists.  A later vl_batch_upload with other reads of
 * the same shapes is accepted as long as it does not have more such entries.
 */
int vl_batch_workspace_bytes(const vl_window_desc* windows, int32_t n_windows, size_t* bytes_out);

/*
 * Host-only (no CUDA call): how the engine will lay out the reads of one window.  The
 * one-hot-expanded read matrix has a column per (position, base); the most frequent base of
 * every position and every other pair that at least max(2, min(64, N/128)) reads carry become
 * dense operand columns (all columns of a position inside one 16-column tile; positions are
 * placed first-fit in ascending order), the remaining entries go to sparse lists.  n_cols_out: dense columns including tile padding; n_sparse_out:
 * sparse entries; col_lb (optional, capacity vl_window_plan_capacity(length)): l * 5 + b per
 * dense column, -1 for padding.  Only mode, n_reads, length and codes of the desc are read.
 */
/* Upper bound of n_cols_out for a window of `length` positions: every tile but the last keeps at
 * least 12 of its 16 columns in use (a position has at most 5), so 16 * (ceil(5 * length / 12) + 1). */
int32_t vl_window_plan_capacity(int32_t length);
int vl_window_plan(const vl_window_desc* window, int32_t* n_cols_out, int64_t* n_sparse_out,
                   int32_t* col_lb);

/*
 * Create a batch on `device` inside a caller-owned device workspace.  `stream` is a
 * cudaStream_t (NULL = the legacy default stream).  Shapes and `codes` are read here (see
 * vl_batch_workspace_bytes); the other host pointers are not.
 */
int vl_batch_create(vl_batch** out, int32_t device, const vl_window_desc* windows,
                    int32_t n_windows, void* dev_workspace, size_t dev_bytes, void* stream);
void vl_batch_destroy(vl_batch* batch);

/*
 * Options, to be set before vl_batch_upload.
 * VL_OPT_SKIP_DEAD (default 1): clusters whose responsibilities are exactly 0 for every read
 * (exp underflow, the normal fate of most of the K clusters with alpha0 << 1) are taken out of
 * the dense work and represented by one shared slot; the results are the ones the dense
 * computation gives.  0 keeps all K clusters in every contraction.
 */
#define VL_OPT_SKIP_DEAD 1
/*
 * VL_OPT_FUSED_CHUNKS (default 1): vl_batch_run enqueues a whole chunk of updates of all running
 * restarts as ONE launch per variant (mode x split haplotype tiles) of one-tile worker CTAs that claim
 * tiles from per-restart scheduling words in device memory, instead of two launches per update; 0
 * restores the per-update launches (vl_batch_step and vl_batch_profile_step always use those).  Same
 * arithmetic, bit-identical results.
 */
#define VL_OPT_FUSED_CHUNKS 3
int vl_batch_set_option(vl_batch* batch, int32_t option, int32_t value);

/* Host -> device copy of every window's inputs and initial state; resets all restarts. */
int vl_batch_upload(vl_batch* batch, const vl_window_desc* windows, int32_t n_windows);

/*
 * Run CAVI for every (window, restart) until each has left its loop (the `while` of
 * cavi.py:120).  Device-resident; returns after the stream has drained.  `n_launches_out`
 * (optional) receives the number of kernels launched.
 */
int vl_batch_run(vl_batch* batch, int64_t* n_launches_out);

/* Run exactly `n_updates` more update()+ELBO iterations on every still-running restart. */
int vl_batch_step(vl_batch* batch, int32_t n_updates, int64_t* n_launches_out);

/* Device -> host copy of results. */
int vl_batch_fetch(vl_batch* batch, vl_window_result* results, int32_t n_windows);

/*
 * Teacher forcing for parity tests: overwrite the running state of one restart with a state
 * of the reference (the keys update_eqs.update reads, update_eqs.py:9-19) and set the loop
 * counter, so that one vl_batch_step reproduces exactly one reference iteration.
 * scalars: [VL_SCALAR_STRIDE], entries MLG0/1, MLT0/1, DSUM_* are read.
 */
int vl_batch_set_state(vl_batch* batch, int32_t window, int32_t restart,
                       const double* mean_cluster, const double* mean_log_pi,
                       const double* scalars, int32_t iter);

/*
 * Finishing contraction (analyze_results.py:65-80): merge the clusters of `restart` with
 * group_of[k] in [0, n_groups) and recompute mean_haplo for the merged responsibilities.
 * merged_cluster_out [N*n_groups] and merged_haplo_out [n_groups*L*B] are host buffers.
 */
int vl_batch_merge_haplo(vl_batch* batch, int32_t window, int32_t restart,
                         const int32_t* group_of, int32_t n_groups,
                         double* merged_cluster_out, double* merged_haplo_out);

/*
 * Like vl_batch_step, but brackets every launch with CUDA events on the launching stream and
 * returns the summed device time per kernel (ms) and the launch counts, arrays of 4:
 * [0] haplotype update, [1] responsibility update incl. the scalar update its last CTA per
 * restart performs (one launch each per update), [2], [3] unused (0).
 * Used by bench.py for the live roofline numbers; not a fast path.
 */
int vl_batch_profile_step(vl_batch* batch, int32_t n_updates, double* ms_out, int64_t* n_out);

/*
 * Per (window, restart) problem, in batch order, four ints as of the last completed update:
 * still running (0/1), 8-cluster tiles its contractions span, clusters in them (live + 1 shared
 * slot for the dead ones), updates done.  out: [4 * n_problems].
 */
int vl_batch_problem_stats(vl_batch* batch, int32_t* out, int32_t n_problems);

/*
 * Host-only (no CUDA call): a window's read file -> compact tensors, the native form of
 * load_and_process_reads (use_quality_scores/preparation.py:24-90; learn_error_params/preparation.py:51-94).
 *
 * vl_reads_parse: `fasta` = the bytes of w-<ref>-<beg>-<end>.reads.fas (one sequence line per record; the
 * buffer must stay alive until vl_reads_free).  unique != 0 collapses identical read strings into the row
 * of their first occurrence.  Returns the number of raw reads, of rows, and the window length.  Fails
 * (vl_last_error) on records of unequal length, multi-line sequences or an empty file; callers fall
 * back to their generic FASTA reader then.
 * vl_reads_fill: codes [n_rows][length] (index in `alphabet`, 255 otherwise), weights [n_rows]
 * (multiplicity), mean_qualities [n_rows][length] = the reference's running mean (q * w + q_new) / (w + 1)
 * in file order, then 0 -> 2 (quality_kind 0: none, 1: int64 [n_raw][length], 2: float64), row_of [n_raw],
 * id_spans [n_raw][2] and seq_spans [n_rows][2] = (offset, length) of the record ids / row sequences inside
 * `fasta` (any of the last three may be NULL).
 */
typedef struct vl_reads vl_reads;
int vl_reads_parse(const char* fasta, size_t n_bytes, int32_t unique, vl_reads** out, int32_t* n_raw_out,
                   int32_t* n_rows_out, int32_t* length_out);
int vl_reads_fill(const vl_reads* reads, const char* alphabet, const void* qualities, int32_t quality_kind,
                  uint8_t* codes, double* weights, double* mean_qualities, int32_t* row_of, int64_t* id_spans,
                  int64_t* seq_spans);
void vl_reads_free(vl_reads* reads);

/*
 * Test hook: the exponential the softmax epilogues use (valid for x <= 0, -inf allowed) applied
 * to n host values on `device`.  It replaces CUDA's exp() in the kernels because FP64 latency
 * bounds the epilogues; tests compare it with the host's exp, including the underflow boundary
 * that decides which responsibilities are exactly 0.
 */
int vl_debug_exp(int32_t device, const double* x, double* out, int32_t n);

/*
 * Development aid: residency of the fused chunk kernel on `device` as the CUDA occupancy calculator
 * reports it: CTAs per SM, shared memory per CTA (static + dynamic) and registers per thread, for
 * kernel class 0..3 (layouts of <= 2 / 4 / 8 / 16 cluster tiles), mode VL_MODE_*, and the deep
 * pipeline (deep != 0) the launcher uses when a chunk has few tiles per update.
 */
int vl_debug_occupancy(int32_t device, int32_t kernel_class, int32_t mode, int32_t deep, int32_t* ctas_per_sm,
                       int32_t* smem_bytes, int32_t* registers);

/*
 * Measurement aid (bench.py's roofline denominator): sustained FP64 tensor throughput of `device` in
 * TFLOP/s, measured now with a register-resident loop of the DMMA m8n8k4 instruction the contraction
 * kernels issue (8 independent accumulators per warp, 16 warps per SM).  The reference has no
 * counterpart (it is a CPU program; SURVEY.md 8(d) asks for the fraction of the FP64 tensor peak).
 */
int vl_debug_fp64_peak(int32_t device, double* tflops);

/*
 * Development aid: per-phase cycle counters, 36 values: [0..15] unused (0), [16..35] tile 0 of
 * restart 0 (slot meanings in vl_kernels.cu).  Fails unless the library was built with -DVL_RS_PROFILE
 * (python -m viloca_b200.build --profile).
 */
int vl_debug_counters(uint64_t* out36, int32_t reset);

/* Milliseconds the last vl_batch_run / vl_batch_step spent on the device (CUDA events). */
double vl_batch_last_device_ms(const vl_batch* batch);

#ifdef __cplusplus
}
#endif
#endif /* VILOCA_B200_H */
