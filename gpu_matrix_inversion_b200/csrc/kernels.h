// Host-side launchers of the sm_100a kernels; every launcher only enqueues on `st`.
// Reference kernel each one replaces is cited at the definition (csrc/*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

// cudaFuncSetAttribute is per device: launchers configure their kernel once on every device they are used on
inline bool first_use_on_device(bool (&flags)[64]) {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (flags[d]) return false;
    flags[d] = true;
    return true;
}

// Per-panel bookkeeping living in device memory (written by kernels, never by the host).
struct PanelState {
    int pos[2 * 128];      // row index of every slot touched by this panel's swaps (first kb = pivot rows)
    int content[2 * 128];  // slot whose ORIGINAL data now sits in this slot
    int m;                 // number of slots in use
    int pad[3];
};

// ---- gj_unblocked.cu : north_star kernels (1) and (2) + rank-1 update, one column per launch trio
void launch_argmax_col(const float *W, long long ld, int n, int col, int row0, u64 *part, int nparts, cudaStream_t st);
void launch_swap_normalize(float *W, long long ld, int n, int r, const u64 *part, int nparts, float *urow,
                           float *ccol, int *piv, int *info, cudaStream_t st);
void launch_rank1_update(float *W, long long ld, int n, int r, const float *urow, const float *ccol, cudaStream_t st);

// ---- gj_panel.cu : panel factorisation (N x kb, kb <= 128)
void launch_panel_step(const float *in, long long ld_in, float *out, long long ld_out, int n, int kb, int t, int k0,
                       const u64 *part_in, int nparts, u64 *part_out, float *CmT, long long ldc, int *piv, float *pv,
                       int *info, PanelState *ps, cudaStream_t st);

// ---- gj_subpanel.cu : panel factorisation v1 (cluster/DSMEM sub-panel kernel + in-panel update)
int subpanel_width(int n);
void panel_set_critical(int on);
bool subpanel_supported(int n);
cudaError_t launch_subpanel(const float *in, long long ld_in, float *out, long long ld_out, int n, int k0, int s0,
                            int sw, float *CmT, long long ldc, int *piv, float *pv, int *info, cudaStream_t st);
void launch_panel_update(const float *in, long long ld_in, float *out, long long ld_out, int n, int k0, int s0, int sw,
                         int wfull, float *CmT, long long ldc, const int *piv, const float *pv, PanelState *ps, int kb,
                         cudaStream_t st);

cudaError_t debug_trace(int on, long long *out128);
cudaError_t debug_trace_pk(int on, long long *out8);
int launch_panel_factor(float *Wv, long long ld, int n, int k0, int kb, float *CmT, long long ldc, int *piv, float *pv,
                        int *info, PanelState *ps, float *P0, float *P1, cudaStream_t st);

// ---- gj_rowblock.cu : row interchanges + row-block recurrence on all non-panel columns
void launch_rowblock(float *W, long long ld, int ncols_pad, int k0, int kb, const float *CmT, long long ldc,
                     const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st);

void launch_rowblock_ex(float *W, long long ld, int ncols_pad, int k0, int kb, int skip_tile, int skip_n, const float *CmT,
                        long long ldc, const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st);

// ---- gj_gemm.cu : trailing update  W[i][j] <- chain_t fma(-CmT[t][i], U[t][j], W[i][j])
void launch_trailing_gemm(float *W, long long ld, int npad, int k0, int kb, const float *CmT, long long ldc,
                          const float *U, long long ldu, cudaStream_t st);

void launch_trailing_gemm_ex(float *W, long long ld, int nrow_tiles, int ncol_tiles, int row_skip, int col_skip, int col_skip_n, int kb,
                             const float *CmT, long long ldc, const float *U, long long ldu, cudaStream_t st);

// ---- gj_gemm_tc.cu : the same trailing update as 3xTF32 on tcgen05 tensor cores (TMEM accumulator); NOT bit-identical
size_t tf32x3_image_bytes(int tiles);
cudaError_t launch_trailing_tf32x3(float *W, long long ld, int nrow_tiles, int ncol_tiles, int row_skip, int col_skip, int col_skip_n,
                                   int kb, const float *CmT, long long ldc, const float *U, long long ldu, float *imgA, float *imgB,
                                   cudaStream_t st);

// ---- gj_probe.cu : O(N^2) randomised estimate of ||A X - I||_F (gate of the 3xTF32 path)
size_t probe_scratch_bytes(int n);
cudaError_t run_probe_residual(const float *A, const float *X, int n, double *scratch, double *out_host, cudaStream_t st);

// ---- gj_finish.cu : deferred column permutation + extraction + isfinite scan
void launch_colperm_build(const int *piv, int n, int *colsrc, cudaStream_t st);
void launch_extract(const float *W, long long ld, int n, const int *colsrc, float *X, int *info, int check,
                    cudaStream_t st);
void launch_extract_rows(const float *W, long long ld, int n, const int *colsrc, float *X, int *info, int check, int row0,
                         int nrows, cudaStream_t st);
void launch_load(const float *A, int n, float *W, long long ld, int npad, cudaStream_t st);
void launch_load_window(const float *A, int n, float *W, long long ld, int npad, int c0, int ncols, cudaStream_t st);

// ---- gj_batched.cu : n <= 128, one CTA per matrix
cudaError_t launch_batched(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st);
cudaError_t launch_batched_pk(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st);
cudaError_t launch_batched_blk(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st);

// ---- generate.cu : synthetic workloads, residual, FFMA peak
void launch_generate(float *A, int n, long long ld, u64 seed, int kind, int col0, int ncols, cudaStream_t st);
void launch_generate_batched(float *A, int n, long long first, long long count, u64 seed0, cudaStream_t st);
cudaError_t run_residual(const float *A, const float *X, int n, double *out_host, cudaStream_t st);
cudaError_t run_residual_f64(const double *A, const double *X, int n, double *out_host, cudaStream_t st);

// FP64 path (gj_f64.cu): unblocked in-place Gauss-Jordan, optional no-pivot mode
struct PivCand;
struct F64Workspace {
    double *W = nullptr, *urow = nullptr, *ccol = nullptr, *io = nullptr;
    double *CT = nullptr, *U = nullptr, *pv = nullptr, *upan = nullptr;   // blocked schedule: multipliers (transposed), snapshots
    double *P[2] = {nullptr, nullptr};                                     // panel ping-pong (n x 64 each)
    PivCand *part = nullptr;
    int *piv = nullptr, *colsrc = nullptr, *info = nullptr;
    int n = 0;
    long long ld = 0;
};
cudaError_t f64_workspace_ensure(F64Workspace &w, int n, bool with_io);
void f64_workspace_free(F64Workspace &w);
int f64_invert_async(F64Workspace &w, const double *A_dev, int n, double *X_dev, int nopivot, int check, cudaStream_t st,
                     cudaEvent_t (*prof_event)());
int f64_invert_blocked_async(F64Workspace &w, const double *A_dev, int n, double *X_dev, int nopivot, int check, cudaStream_t st,
                             cudaEvent_t (*prof_event)());
cudaError_t run_ffma_peak(double *tflops, cudaStream_t st);
