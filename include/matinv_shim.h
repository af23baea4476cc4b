/*
 * matinv_shim.h -- C-ABI boundary of the B200 Gauss-Jordan inverter (libmatinv32.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  It replaces the
 * OpenCL platform / context / program / queue / buffer code of the reference's host function
 * (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp, abbreviated LIB below) and
 * is what mat_inv_32.cpp (the C++ `matrix_inv_32`), the Python ctypes host
 * (gpu_matrix_inversion_b200/) and bench.py bind.  INTEGRATION.md shows the reference-side
 * binding.
 *
 * Conventions (all entry points):
 *   - matrices are row-major, element (i,j) at i*n + j, FP32       (LIB:183, LIB:201)
 *   - return 0 = ok, 1 = singular / non-finite (the C++ layer returns {}),
 *     < 0 = error (MATINV_E_*); matinv_last_error() holds the text (thread-local)
 *   - never throws, never aborts; no CPU fallback: without a CUDA device every compute entry
 *     returns MATINV_E_NODEVICE
 *   - 64-bit indexing internally (the reference overflows `int` at n >= 32768, LIB:232, LIB:70)
 */
#ifndef MATINV_SHIM_H
#define MATINV_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

#define MATINV_OK 0
#define MATINV_SINGULAR 1
#define MATINV_E_INVALID (-1)  /* bad argument                                   */
#define MATINV_E_NODEVICE (-2) /* no CUDA device (reference: UB at LIB:239-240)  */
#define MATINV_E_CUDA (-3)     /* CUDA runtime error, text in matinv_last_error  */
#define MATINV_E_UNSUPPORTED (-4)

/* flags */
#define MATINV_FLAG_TF32X3 1    /* trailing update as 3xTF32 on tcgen05 tensor cores; NOT bit-exact, accepted only
                                   if the residual estimate passes MATINV_TF32X3_GATE, else the FP32 path is rerun */
#define MATINV_FLAG_UNBLOCKED 2 /* force the unblocked 3-kernel-per-column path (parity checks)  */
#define MATINV_FLAG_VERBOSE 4   /* print the reference's two stdout lines (LIB:385-386)          */
#define MATINV_FLAG_NOCHECK 8   /* skip the final isfinite scan (kept off the hot path timing)   */
#define MATINV_FLAG_NOPIVOT 16  /* FP64 entries only: pivot = diagonal entry, no row interchange    */

/* Gate of MATINV_FLAG_TF32X3: estimate of ||A X - I||_F / (n ||A||_F ||X||_F) must not exceed this (north_star bound). */
#define MATINV_TF32X3_GATE 1e-5
/* ... and the same estimate times sqrt(n), which does not shrink with the order (the FP32 SIMT path measures ~3e-9), must
 * not exceed this: without it the 1e-5 bound only rejects NaN and total garbage at large n (an unrelated X scores ~n^-1.5). */
#define MATINV_TF32X3_GATE_SCALED 1e-7

/* Replaces cl::Platform::get / getDevices (LIB:239-244).  Number of usable CUDA devices, 0 if none. */
int matinv_device_count(void);

/* Text of the last error on this thread ("" if none). */
const char *matinv_last_error(void);

/* Frees the process-wide context (streams, workspaces).  Replaces buffers.clear() (LIB:388). */
void matinv_shutdown(void);

/* Whole `matrix_inv_32` body below the argument checks (LIB:219-389): H2D, N Gauss-Jordan steps
 * with partial pivoting, extraction, D2H.  A_host/X_host: n*n floats (may alias); piv_host:
 * n ints or NULL, piv[r] = row chosen as pivot at step r (arg max |.| over rows >= r, lowest
 * index on ties). */
int matinv_invert_f32(const float *A_host, int n, float *X_host, int *piv_host, int flags);

/* Same with device pointers on the current device: the "Tempo Computazione" window (LIB:316-365)
 * plus extraction (LIB:369-376), no PCIe traffic.  Work is enqueued on `stream` (a cudaStream_t,
 * NULL = default stream); the call returns after the stream has drained because the status word
 * is read back.  A_dev and X_dev may alias. */
int matinv_invert_f32_dev(const float *A_dev, int n, float *X_dev, int *piv_dev, void *stream, int flags);

/* Batched small-n path (n <= 128), one CTA per matrix.  The reference has no batched entry: this
 * is `for b: matrix_inv_32(A[b], n)` in one launch.  A/X: batch*n*n floats; info: per-matrix
 * status (0 ok, r+1 zero/non-finite pivot at step r, -1 non-finite inverse) or NULL.
 * Returns 0 when every matrix inverted, 1 when at least one is singular, < 0 on error. */
int matinv_invert_batched_f32(const float *A_host, int n, long long batch, float *X_host, int *info_host,
                              int flags);
int matinv_invert_batched_f32_dev(const float *A_dev, int n, long long batch, float *X_dev, int *info_dev,
                                  void *stream, int flags);

/* ---- single-call multi-GPU entries (one process, one host thread per GPU, NCCL) --------------
 * SURVEY.md s.8(b)/(e).  The reference is single-device (LIB:239-250); these are what `matrix_inv_32` reaches when
 * MATINV_NGPU > 1.  ngpu <= 0 selects MATINV_NGPU or, if unset, every visible device.
 *
 * matinv_invert_sharded_f32: one n x n matrix, columns dealt block-cyclically (blocks of 128, nb must be 0 or 128) to ngpu
 * GPUs; per block step the owner factors the panel and ncclBroadcast ships pivots + multipliers; one grouped
 * ncclSend/ncclRecv applies the deferred column permutation.  Result and pivot sequence are bit-identical to
 * matinv_invert_f32.  Returns 0 / 1 (singular or non-finite) / < 0; MATINV_E_UNSUPPORTED when ngpu > 1 and NCCL cannot be
 * loaded (libnccl.so.2 is opened with dlopen at first use). */
int matinv_invert_sharded_f32(const float *A_host, int n, float *X_host, int *piv_host, int ngpu, int nb, int flags);
/* Same schedule on the synthetic workload generated on the devices (no host matrix): timing aid for orders whose host
 * copy would dominate.  compute_ms (may be NULL) = factorisation + column exchange on rank 0, CUDA events. */
int matinv_sharded_synthetic_f32(int n, unsigned long long seed, int kind, int ngpu, int *piv_host, double *compute_ms);
/* Batched small-n path split by matrix index over ngpu GPUs (contiguous index ranges, no communication).
 * matinv_invert_batched_f32 takes this route when MATINV_NGPU > 1. */
int matinv_invert_batched_f32_ngpu(const float *A_host, int n, long long batch, float *X_host, int *info_host, int ngpu,
                                   int flags);
/* NCCL version code of the library the multi-GPU entries would use (e.g. 22703), 0 if none can be loaded. */
int matinv_nccl_version(void);

/* ---- column-sharded single inversion: per-rank primitives, one process per GPU ---------------
 * The host (Python + torch.distributed, or any MPI-like launcher) owns the exchange step; these
 * calls own the math.  Columns are dealt block-cyclically in blocks of 128: global column block J
 * lives on rank J % world.  See DESIGN.md "multi-GPU". */
/* Thread-safety: handles are independent; one host thread per device may drive its own handle concurrently (that is what the
 * single-call entries above do).  Do not run MATINV_FLAG_TF32X3 inversions concurrently with them: that mode switches the
 * panel kernels to their critical-path shapes process-wide (results are unaffected -- the shapes are bit-identical -- but the
 * tuning is not). */
typedef struct matinv_shard matinv_shard_t;
/* panel message: what the owner of block J broadcasts.  Size in bytes for a given n. */
long long matinv_shard_panel_bytes(int n);
int matinv_shard_create(int n, int rank, int world, matinv_shard_t **out);
void matinv_shard_destroy(matinv_shard_t *s);
/* local storage: n_pad x local_cols floats, row-major with leading dimension local_ld */
float *matinv_shard_local(matinv_shard_t *s, long long *local_cols, long long *local_ld);
/* copy global column block J (n x 128, leading dimension ld, host or device memory) into / out of the shard;
 * only valid on the owner of J */
int matinv_shard_set_block(matinv_shard_t *s, int J, const float *src, long long ld, void *stream);
int matinv_shard_get_block(matinv_shard_t *s, int J, float *dst, long long ld, void *stream);
/* fill the local columns with the synthetic workload (same bits as the unsharded generator) */
int matinv_shard_generate(matinv_shard_t *s, unsigned long long seed, int kind, void *stream);
/* owner only: factor panel J in place and pack the message into panel_dev */
int matinv_shard_factor(matinv_shard_t *s, int J, void *panel_dev, void *stream);
/* every rank: apply panel J's swaps + row-block recurrence + trailing update to the local columns */
int matinv_shard_apply(matinv_shard_t *s, int J, const void *panel_dev, void *stream);
/* look-ahead split of the same update: mode 1 = only global block `block` (local), mode 2 = all local columns but it */
int matinv_shard_apply_ex(matinv_shard_t *s, int J, const void *panel_dev, void *stream, int mode, int block);
/* status word (0 ok / r+1 / -1) and the n pivot rows; the deferred column permutation X[:, j] = M[:, colsrc[j]]
 * follows from piv and is applied by the host, which moves columns between shards accordingly */
int matinv_shard_status(matinv_shard_t *s, int *info_host, int *piv_host, void *stream);

/* ---- synthetic workloads and checks on the device (bench + tests; same bits as oracle/) ------
 * kind 0: U[0,100) counter-based; kind 1: diagonally dominant.  Fills A_dev[i*ld + (j-col0)] for
 * j in [col0, col0+ncols) of the n x n matrix with the given seed. */
int matinv_generate_f32_dev(float *A_dev, int n, long long ld, unsigned long long seed, int kind, int col0,
                            int ncols, void *stream);
/* Batched generator: matrix b (global index first+b) uses seed seed0 + first + b. */
int matinv_generate_batched_f32_dev(float *A_dev, int n, long long first, long long count,
                                    unsigned long long seed0, void *stream);
/* ||A X - I||_F^2, ||A||_F^2, ||X||_F^2 in FP64 on the device (verification GEMM, replaces
 * SOL/matrix_multiply.cpp:15-212).  out_host[3]. */
int matinv_residual_f32_dev(const float *A_dev, const float *X_dev, int n, double *out_host, void *stream);

/* ---- FP64 entry points -------------------------------------------
 * Device side of the development copy's double-precision functions
 *   matrix_inversion_FP64        (/root/reference/matrix_inv_solution/matrix_inversion_solution/matrix_inversion/matrix_inversion_FP64.cpp:13)
 *   matrix_inversion_no_pivots   (.../matrix_inversion_no_pivots.cpp:10)   -> flags |= MATINV_FLAG_NOPIVOT
 * Same conventions as the FP32 entries (row-major, 0 ok / 1 singular or non-finite / < 0 error, piv optional).
 * Blocked schedule with 64-column panels; MATINV_FLAG_UNBLOCKED selects the per-column schedule (cross-check). */
int matinv_invert_f64(const double *A_host, int n, double *X_host, int *piv_host, int flags);
int matinv_invert_f64_dev(const double *A_dev, int n, double *X_dev, int *piv_dev, void *stream, int flags);
/* FP64-input twin of matinv_residual_f32_dev with a fourth output: out_host[4] = ||AX-I||_F^2, ||A||_F^2, ||X||_F^2,
 * ||AX||_F^2.  The C++ `matrix_multiply` (SOL/matrix_multiply.cpp:15-212: sqrt(n) - ||A B||_F) is built on it. */
int matinv_residual_f64_dev(const double *A_dev, const double *X_dev, int n, double *out_host, void *stream);
/* Same with host pointers (copies both operands to the device first). */
int matinv_host_defect_f64(const double *A_host, const double *B_host, int n, double *out_host);

/* Seconds spent by the last matinv_invert_f32 on this thread: total (H2D+compute+D2H, the
 * reference's "Tempo Totale") and compute ("Tempo Computazione").  Returns 0 if available. */
int matinv_last_timing(double *total_s, double *compute_s);

/* Phase split of the last matinv_invert_f32 on this thread, the phases the reference's instrumented copy records
 * (SOL/FP32_bench.cpp:256-443, Res.times): out5 = {setup (streams / staging buffers), H2D, factorisation, extraction + D2H
 * (overlapped) + status read-back, total}, seconds; device phases from CUDA events.  The same phases are NVTX ranges
 * ("matinv_invert_f32", "H2D", ...) for Nsight Systems.  Returns 0 if available (not for MATINV_FLAG_TF32X3 calls). */
int matinv_last_phases(double *out5);

/* Host logic of the pipelined upload of matinv_invert_f32 (pinned buffers, n >= 8192), exposed so that its
 * invariants can be tested without a GPU: the column windows (start column c0, width ncols, both multiples of 128, tiling
 * [0, npad)), the panel at which each window joins (act; window w must be active before its first block becomes the
 * look-ahead block: act[w] <= c0[w] / 128 - 1) and the number of message slots kept.  Arrays of 8.  Returns 1 if this order is
 * pipelined, 0 if not (then *nwin = 0), < 0 on a bad argument.  Makes no CUDA call. */
int matinv_debug_pipeline_plan(int n, int *nwin, int *c0_8, int *ncols_8, int *act_8, int *ring_slots);

/* Profiling hooks for bench.py.  With profiling enabled the shim brackets every trailing-update
 * (GEMM) launch with CUDA events on the launching stream.  matinv_profile_read returns, for the
 * calls made since the last matinv_profile_enable: summed GEMM time (ms), number of GEMM launches,
 * algorithmic flops of those launches, and the number of kernels launched in total. */
void matinv_profile_enable(int on);
int matinv_profile_read(double *gemm_ms, long long *gemm_launches, double *gemm_flops, long long *all_launches);

/* O(n^2) randomised estimate of the same three numbers (4 Rademacher probe vectors: E ||(A X - I) v||^2 = ||A X - I||_F^2);
 * out_host[0] is an estimate, out_host[1..2] are exact.  This is what gates MATINV_FLAG_TF32X3. */
int matinv_probe_residual_f32_dev(const float *A_dev, const float *X_dev, int n, double *out_host, void *stream);

/* The gate of MATINV_FLAG_TF32X3 applied to a caller-provided pair (test hook / diagnostics): runs the O(n^2) probe and the
 * acceptance rule the gated inversion uses.  est_out[2] (may be NULL) = {estimate, estimate * sqrt(n)}.  Returns 1 =
 * accepted, 0 = rejected (NaN included), < 0 = error. */
int matinv_tf32x3_gate_dev(const float *A_dev, const float *X_dev, int n, double *est_out, void *stream);

/* MATINV_FLAG_TF32X3 bookkeeping: residual estimate of the last gated inversion (-1 if none), whether it fell back to
 * the FP32 SIMT schedule, and the totals since the library was loaded.  Any pointer may be NULL. */
int matinv_tf32x3_status(double *last_estimate, int *last_fallback, long long *inversions, long long *fallbacks);

/* Test hook: ONE trailing update  W[i][j] -= sum_t CmT[t][i] U[t][j]  (128 steps) on an npad x npad matrix whose pivot
 * rows / panel columns are block k0/128; CmT and U are 128 x npad with leading dimension ld.  mode 0 = FP32 SIMT kernel
 * (gj_gemm.cu), mode 1 = 3xTF32 tcgen05 kernel (gj_gemm_tc.cu).  The update is applied `reps` times back to back; avg_ms (may
 * be NULL) receives the CUDA-event time per application.  Synchronises the stream. */
int matinv_debug_trailing_update(float *W_dev, long long ld, int npad, int k0, const float *CmT_dev, const float *U_dev,
                                 int mode, int reps, double *avg_ms, void *stream);

/* Tuning aid: switch the in-kernel timeline of the panel kernels on/off and read the SM-clock stamps of
 * the last launch (128 slots; see gj_subpanel.cu).  out128 may be NULL. */
int matinv_debug_trace(int on, long long *out128);

/* FFMA throughput micro-benchmark (TFLOP/s) -- the FP32 SIMT roofline denominator, measured live. */
int matinv_ffma_peak_tflops(double *tflops_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MATINV_SHIM_H */
