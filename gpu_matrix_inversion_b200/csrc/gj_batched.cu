// Batched small-n Gauss-Jordan: one CTA per matrix, matrix resident in shared memory for the whole
// inversion (north_star kernel 4).  The reference has no batched entry -- this is
// `for b: matrix_inv_32(A[b], n)` (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:11-395)
// collapsed into one launch; per matrix the arithmetic is the in-place form of SURVEY.md
// Appendix A.3, so results are bit-identical to the large-n path and to the oracle.
//
// v0 layout: a[n][n+1] floats in shared memory, 256 threads, three barriers per column.
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256)
batched_smem_kernel(const float *__restrict__ A, int n, long long batch, float *__restrict__ X,
                    int *__restrict__ info) {
    extern __shared__ float sm[];
    const int ldm = n + 1;
    float *a = sm;                        // n * (n+1)
    float *urow = a + n * ldm;            // n
    float *ccol = urow + n;               // n
    int *piv = reinterpret_cast<int *>(ccol + n);  // n
    __shared__ u64 skey[8];
    __shared__ int sinfo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float *Ab = A + b * (long long)n * n;
        float *Xb = X + b * (long long)n * n;
        for (int e = tid; e < n * n; e += 256) a[(e / n) * ldm + (e % n)] = Ab[e];
        if (tid == 0) sinfo = 0;
        __syncthreads();

        for (int r = 0; r < n; r++) {
            // (1) pivot search: rows r..n-1 of column r
            u64 k = 0;
            for (int i = r + tid; i < n; i += 256) {
                const u64 kk = gj_key(a[i * ldm + r], i, i == r);
                k = kk > k ? kk : k;
            }
            k = warp_max_u64(k);
            if (lane == 0) skey[warp] = k;
            __syncthreads();
            u64 best = skey[0];
#pragma unroll
            for (int w = 1; w < 8; w++) best = skey[w] > best ? skey[w] : best;
            const int p = gj_key_row(best);
            const float v = gj_key_value(best);
            if (tid == 0) {
                piv[r] = p;
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
            }
            // (2) swap + normalise: read phase
            float rr = 0.f, pp = 0.f, cc = 0.f;
            if (tid < n) {
                rr = a[r * ldm + tid];
                pp = a[p * ldm + tid];
                cc = a[tid * ldm + r];
            }
            const float arr = a[r * ldm + r];
            __syncthreads();
            if (tid < n) {
                const float u = (tid == r) ? 1.0f / v : pp / v;
                urow[tid] = u;
                a[r * ldm + tid] = u;
                if (p != r) a[p * ldm + tid] = rr;
                ccol[tid] = (tid == r) ? 0.0f : ((tid == p) ? arr : cc);
            }
            __syncthreads();
            // (3) rank-1 update of every row but r
            for (int e = tid; e < n * n; e += 256) {
                const int i = e / n, j = e - i * n;
                if (i == r) continue;
                const float c = ccol[i], u = urow[j];
                a[i * ldm + j] = (j == r) ? fmaf(-c, u, 0.0f) : gj_elim(a[i * ldm + j], c, u);
            }
            __syncthreads();
        }
        // deferred column permutation (see gj_finish.cu:colperm_kernel)
        int *colsrc = reinterpret_cast<int *>(urow);
        __syncthreads();
        if (tid < n) {
            int c = tid;
            for (int r = 0; r < n; r++) {
                if (c < r) break;
                const int p = piv[r];
                if (c == r) c = p;
                else if (c == p) c = r;
            }
            colsrc[tid] = c;
        }
        __syncthreads();
        bool bad = false;
        for (int e = tid; e < n * n; e += 256) {
            const int i = e / n, j = e - i * n;
            const float x = a[i * ldm + colsrc[j]];
            bad |= !isfinite(x);
            Xb[e] = x;
        }
        const int anybad = __syncthreads_or(bad);
        if (tid == 0 && info) info[b] = sinfo ? sinfo : (anybad ? -1 : 0);
        __syncthreads();
    }
}

cudaError_t launch_batched(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st) {
    const size_t smem = ((size_t)n * (n + 1) + 3 * (size_t)n) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(batched_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 129 * 4 + 3 * 128 * 4);
        configured = true;
    }
    long long grid = batch;
    if (grid > (1ll << 20)) grid = 1ll << 20;
    batched_smem_kernel<<<(unsigned)grid, 256, smem, st>>>(A, n, batch, X, info);
    return cudaGetLastError();
}
