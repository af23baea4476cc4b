// Load / extract passes around the in-place factorisation.
//
//   load_kernel      A (n x n, contiguous) -> W (npad x npad workspace, zero padded).  Takes the place of
//                    makeAugmentedMatrix (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:177-192):
//                    the identity half is never materialised by the in-place form.
//   colperm_kernel   net effect of the deferred column swaps `for r = n-1..0: swap cols r, piv[r]`
//                     as a gather list colsrc[j].
//   extract_kernel   X[i][j] = W[i][colsrc[j]] + isfinite scan.  Takes the place of getInvertedMatrix
//                    (:195-203) and of the host-side identity check of
//                    matrix_inv_solution/.../matrix_inversion_FP32.cpp:814-835 (non-finite => {}).
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256) load_kernel(const float *__restrict__ A, int n, float *__restrict__ W,
                                                   long long ld, int npad) {
    const long long i = blockIdx.x;   // rows on grid.x: grid.y is limited to 65535
    const int j = blockIdx.y * 256 + threadIdx.x;
    if (j >= npad) return;
    W[i * ld + j] = (i < n && j < n) ? A[i * (long long)n + j] : 0.0f;
}

// X = M * P_{n-1} ... P_0.  Column j of X is column q(j) of M with q(j) = pi_{n-1}(...pi_0(j)),
// pi_r the transposition (r, piv[r]).  Once the walker sits on a column < r it can never move again
// (later transpositions only involve columns >= r), so a CTA stops as soon as all of its walkers are
// settled.  The pivots are staged through shared memory in chunks so that the walk itself is a chain of
// compares fed by independent LDS (the first version read piv[r] from global memory inside the chain:
// 125 cycles per transposition, 1 ms at n = 16384).
#define COLPERM_CHUNK 1024
__global__ void __launch_bounds__(256) colperm_kernel(const int *__restrict__ piv, int n, int *__restrict__ colsrc) {
    __shared__ int sp[COLPERM_CHUNK];
    const int j = blockIdx.x * 256 + threadIdx.x;
    int c = j;
    for (int r0 = 0; r0 < n; r0 += COLPERM_CHUNK) {
        const int len = min(COLPERM_CHUNK, n - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += 256) sp[i] = piv[r0 + i];
        __syncthreads();
        if (c >= r0) {
#pragma unroll 8
            for (int i = 0; i < len; i++) {
                const int r = r0 + i, p = sp[i];
                c = (c == r) ? p : ((c == p) ? r : c);
            }
        }
        if (!__syncthreads_or(c >= r0 + len)) break;
    }
    if (j < n) colsrc[j] = c;
}

// One CTA per row; the row is staged in shared memory so both the read of W and the write of X are
// coalesced.  Falls back to a direct gather when the row does not fit (n > 48K).
__global__ void __launch_bounds__(512) extract_kernel(const float *__restrict__ W, long long ld, int n,
                                                      const int *__restrict__ colsrc, float *__restrict__ X,
                                                      int *__restrict__ info, int check, int staged, int row0) {
    extern __shared__ float srow[];
    const long long i = (long long)blockIdx.x + row0;
    const float *wr = W + i * ld;
    float *xr = X + i * (long long)n;
    bool bad = false;
    if (staged) {
        for (int j = threadIdx.x; j < n; j += 512) srow[j] = wr[j];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += 512) {
            const float v = srow[colsrc[j]];
            bad |= !isfinite(v);
            xr[j] = v;
        }
    } else {
        for (int j = threadIdx.x; j < n; j += 512) {
            const float v = wr[colsrc[j]];
            bad |= !isfinite(v);
            xr[j] = v;
        }
    }
    if (check && __syncthreads_or(bad) && threadIdx.x == 0) atomicCAS(info, 0, -1);
}

// columns [c0, c0 + ncols) only (c0 + ncols <= npad): the host entry uploads A in column windows and loads each on arrival
__global__ void __launch_bounds__(256) load_window_kernel(const float *__restrict__ A, int n, float *__restrict__ W,
                                                          long long ld, int c0, int ncols) {
    const long long i = blockIdx.x;
    const int jj = blockIdx.y * 256 + threadIdx.x;
    if (jj >= ncols) return;
    const int j = c0 + jj;
    W[i * ld + j] = (i < n && j < n) ? A[i * (long long)n + j] : 0.0f;
}

void launch_load_window(const float *A, int n, float *W, long long ld, int npad, int c0, int ncols, cudaStream_t st) {
    dim3 grid(npad, (ncols + 255) / 256);
    load_window_kernel<<<grid, 256, 0, st>>>(A, n, W, ld, c0, ncols);
}

void launch_load(const float *A, int n, float *W, long long ld, int npad, cudaStream_t st) {
    dim3 grid(npad, (npad + 255) / 256);
    load_kernel<<<grid, 256, 0, st>>>(A, n, W, ld, npad);
}

void launch_colperm_build(const int *piv, int n, int *colsrc, cudaStream_t st) {
    colperm_kernel<<<(n + 255) / 256, 256, 0, st>>>(piv, n, colsrc);
}

// rows [row0, row0 + nrows) of X (X points at the full n x n output)
void launch_extract_rows(const float *W, long long ld, int n, const int *colsrc, float *X, int *info, int check, int row0,
                         int nrows, cudaStream_t st) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    const size_t bytes = (size_t)n * sizeof(float);
    const int staged = bytes <= 200 * 1024;
    extract_kernel<<<nrows, 512, staged ? bytes : 0, st>>>(W, ld, n, colsrc, X, info, check, staged, row0);
}

void launch_extract(const float *W, long long ld, int n, const int *colsrc, float *X, int *info, int check,
                    cudaStream_t st) {
    launch_extract_rows(W, ld, n, colsrc, X, info, check, 0, n, st);
}
