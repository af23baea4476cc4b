// Unblocked in-place Gauss-Jordan: three launches per column.
//
//   argmax_col_kernel       (1) warp-shuffle arg max |W[i][col]| over rows i >= row0
//                               replaces maxPivotKernel + finalMaxPivotKernel
//                               (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:61-132)
//   swap_normalize_kernel   (2) fused row swap + pivot-row normalisation (true division)
//                               replaces pivotElementsKernel (:154-173) + fixRowKernel (:138-150)
//   rank1_update_kernel         W[i][j] <- fma(-c_i, u_j, W[i][j]); replaces fixColumnKernel (:13-57)
//
// This path is the GPU-side cross-check of the blocked path and serves n <= 256 directly.
#include "common.cuh"
#include "kernels.h"

// One partial per MATINV_RB rows; partial g covers rows [g*RB, (g+1)*RB) intersected with
// [row0, n).  Rows outside contribute the empty key 0.
__global__ void __launch_bounds__(MATINV_RB) argmax_col_kernel(const float *__restrict__ W, long long ld, int n,
                                                               int col, int row0, u64 *__restrict__ part) {
    __shared__ u64 sm[MATINV_RB / 32];
    const int i = blockIdx.x * MATINV_RB + threadIdx.x;
    u64 k = 0;
    if (i >= row0 && i < n) k = gj_key(W[(long long)i * ld + col], i, i == row0);
    k = warp_max_u64(k);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 b = sm[0];
#pragma unroll
        for (int w = 1; w < MATINV_RB / 32; w++) b = sm[w] > b ? sm[w] : b;
        part[blockIdx.x] = b;
    }
}

// Every CTA reduces the partials redundantly (nparts*8 bytes, L2 resident) -- no grid sync.
__device__ __forceinline__ u64 reduce_partials(const u64 *__restrict__ part, int nparts, u64 *sm /*[32]*/) {
    u64 k = 0;
    for (int g = threadIdx.x; g < nparts; g += blockDim.x) {
        const u64 o = part[g];
        k = o > k ? o : k;
    }
    k = warp_max_u64(k);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = k;
    __syncthreads();
    u64 b = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; w++) b = sm[w] > b ? sm[w] : b;
    return b;
}

// grid: ceil(max(n,1)/256) CTAs; thread x handles column j = x of the two rows AND row i = x of
// the multiplier column.  urow[j] = normalised pivot row (urow[r] = 1/v); ccol[i] = multiplier of
// row i after the swap (ccol[r] = 0).
__global__ void __launch_bounds__(256) swap_normalize_kernel(float *__restrict__ W, long long ld, int n, int r,
                                                             const u64 *__restrict__ part, int nparts,
                                                             float *__restrict__ urow, float *__restrict__ ccol,
                                                             int *__restrict__ piv, int *__restrict__ info) {
    __shared__ u64 sm[32];
    const u64 best = reduce_partials(part, nparts, sm);
    const int p = gj_key_row(best);
    const float v = gj_key_value(best);  // carried by the key: no re-read of W[p][r] (it is rewritten below)
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x == 0) {
        piv[r] = p;
        if (gj_bad_pivot(v) && *info == 0) *info = r + 1;
    }
    if (x >= n) return;
    // column part first: reads of rows != r,p never race with the row part below
    if (x != r && x != p) ccol[x] = W[(long long)x * ld + r];
    // row part
    const float rr = W[(long long)r * ld + x];
    const float pp = W[(long long)p * ld + x];
    const float u = (x == r) ? 1.0f / v : pp / v;
    urow[x] = u;
    W[(long long)r * ld + x] = u;
    if (p != r) W[(long long)p * ld + x] = rr;
    if (x == r) {
        ccol[r] = 0.0f;
        if (p != r) ccol[p] = rr;
    }
}

// grid (ceil(n/128), ceil(n/8)), block (32, 8): each thread 4 consecutive columns of one row.
__global__ void __launch_bounds__(256) rank1_update_kernel(float *__restrict__ W, long long ld, int n, int r,
                                                           const float *__restrict__ urow,
                                                           const float *__restrict__ ccol) {
    const int i = blockIdx.y * 8 + threadIdx.y;
    const int j0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    if (i >= n || j0 >= n || i == r) return;
    const float c = ccol[i];
    float *row = W + (long long)i * ld;
    if (j0 + 3 < n && (ld & 3) == 0) {
        float4 a = *reinterpret_cast<float4 *>(row + j0);
        const float4 u = *reinterpret_cast<const float4 *>(urow + j0);
        a.x = (j0 + 0 == r) ? fmaf(-c, u.x, 0.0f) : gj_elim(a.x, c, u.x);
        a.y = (j0 + 1 == r) ? fmaf(-c, u.y, 0.0f) : gj_elim(a.y, c, u.y);
        a.z = (j0 + 2 == r) ? fmaf(-c, u.z, 0.0f) : gj_elim(a.z, c, u.z);
        a.w = (j0 + 3 == r) ? fmaf(-c, u.w, 0.0f) : gj_elim(a.w, c, u.w);
        *reinterpret_cast<float4 *>(row + j0) = a;
    } else {
        for (int j = j0; j < n && j < j0 + 4; j++) {
            const float u = urow[j];
            row[j] = (j == r) ? fmaf(-c, u, 0.0f) : gj_elim(row[j], c, u);
        }
    }
}

void launch_argmax_col(const float *W, long long ld, int n, int col, int row0, u64 *part, int nparts,
                       cudaStream_t st) {
    argmax_col_kernel<<<nparts, MATINV_RB, 0, st>>>(W, ld, n, col, row0, part);
}

void launch_swap_normalize(float *W, long long ld, int n, int r, const u64 *part, int nparts, float *urow,
                           float *ccol, int *piv, int *info, cudaStream_t st) {
    swap_normalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(W, ld, n, r, part, nparts, urow, ccol, piv, info);
}

void launch_rank1_update(float *W, long long ld, int n, int r, const float *urow, const float *ccol,
                         cudaStream_t st) {
    dim3 grid((n + 127) / 128, (n + 7) / 8), block(32, 8);
    rank1_update_kernel<<<grid, block, 0, st>>>(W, ld, n, r, urow, ccol);
}
