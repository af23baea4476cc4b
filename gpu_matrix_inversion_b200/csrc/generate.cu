// Synthetic workloads, on-device verification and the FFMA roofline probe.
//
//   generate_kernel     counter-based U[0,100) FP32 (distribution of matrix_inv_pyopencl.py:17 /
//                       matrix_inv_numpy.py:40 in /root/reference), bit-identical to oracle/gj_oracle.c
//   residual_kernel     ||A X - I||_F^2, ||A||_F^2, ||X||_F^2 with FP64 accumulation -- the verification
//                       GEMM of matrix_inv_solution/.../matrix_multiply.cpp:15-212, on the device
//   ffma_peak_kernel    dependent-chain FFMA throughput: the measured FP32 SIMT roofline denominator
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256) generate_kernel(float *__restrict__ A, int n, long long ld, u64 seed,
                                                       int col0, int ncols) {
    const long long i = blockIdx.x;
    const int jl = blockIdx.y * 256 + threadIdx.x;
    if (jl >= ncols) return;
    const int j = col0 + jl;
    A[i * ld + jl] = gj_u100(seed, (u64)i * (u64)n + (u64)j);
}

// Diagonal of the diagonally-dominant family: one warp per row, but the FP32 sum must run in j
// order to match the oracle, so each row is summed sequentially by one thread over values that
// are regenerated on the fly (no dependence on which columns are local).
__global__ void __launch_bounds__(128) diagdom_kernel(float *__restrict__ A, int n, long long ld, u64 seed, int col0,
                                                      int ncols) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n || i < col0 || i >= col0 + ncols) return;
    float s = 0.0f;
    for (int j = 0; j < n; j++)
        if (j != i) s = __fadd_rn(s, gj_u100(seed, (u64)i * (u64)n + (u64)j));
    A[(long long)i * ld + (i - col0)] = __fadd_rn(__fadd_rn(s, gj_u100(seed, (u64)i * (u64)n + (u64)i)), 1.0f);
}

__global__ void __launch_bounds__(256) generate_batched_kernel(float *__restrict__ A, int n, long long first,
                                                               long long count, u64 seed0) {
    const long long nn = (long long)n * n;
    const long long total = count * nn;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const long long b = e / nn;
        A[e] = gj_u100(seed0 + (u64)(first + b), (u64)(e - b * nn));
    }
}

// 64x64 tile of R = A X, FP64 accumulate, 256 threads x (4x4).
template <typename T>
__global__ void __launch_bounds__(256) residual_kernel(const T *__restrict__ A, const T *__restrict__ X, int n,
                                                       double *__restrict__ out) {
    __shared__ T sa[16][64 + 1];
    __shared__ T sx[16][64];
    __shared__ double red[4][8];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    double acc[4][4] = {};
    double a2 = 0.0, x2 = 0.0;
    for (int k0 = 0; k0 < n; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int ii = e >> 4, kk = e & 15;
            const T v = (i0 + ii < n && k0 + kk < n) ? A[(long long)(i0 + ii) * n + k0 + kk] : (T)0;
            sa[kk][ii] = v;
            if (blockIdx.x == 0) a2 += (double)v * (double)v;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = e >> 6, jj = e & 63;
            const T v = (k0 + kk < n && j0 + jj < n) ? X[(long long)(k0 + kk) * n + j0 + jj] : (T)0;
            sx[kk][jj] = v;
            if (blockIdx.y == 0) x2 += (double)v * (double)v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            double av[4], xv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { av[q] = (double)sa[kk][ty * 4 + q]; xv[q] = (double)sx[kk][tx * 4 + q]; }
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int w = 0; w < 4; w++) acc[q][w] = fma(av[q], xv[w], acc[q][w]);
        }
        __syncthreads();
    }
    double r2 = 0.0, p2 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int i = i0 + ty * 4 + q, j = j0 + tx * 4 + w;
            if (i < n && j < n) {
                const double d = acc[q][w] - (i == j ? 1.0 : 0.0);
                r2 += d * d;
                p2 += acc[q][w] * acc[q][w];
            }
        }
    double vals[4] = {r2, a2, x2, p2};   // out[3] = ||A X||_F^2 (the reference's "Frobenius defect" is sqrt(n) - its root)
#pragma unroll
    for (int q = 0; q < 4; q++) {
        double v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}

// Probe 1: 8 independent FMA chains per thread.  Probe 2: the register-tile outer product of the GEMM inner
// loop (64 FFMA on 8+8 operands, no loads).  Results are written so the loops are not dead code.
__global__ void __launch_bounds__(1024) ffma_peak_kernel(float *out, int iters, float seed) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
          a7 = seed + 7;
    const float m = 0.999f + seed * 1e-9f, c = 1e-3f * (threadIdx.x & 3);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678f) out[blockIdx.x * 1024 + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256, 2) ffma_tile_kernel(float *out, const float *__restrict__ in, int iters) {
    float a[8], b[8], acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[(threadIdx.x + i) & 255]; b[i] = in[(threadIdx.x + 8 + i) & 255]; }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = (float)(i * 8 + j);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = fmaf(-a[i], b[j], acc[i][j]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += acc[i][j];
    if (s == 12345.678f) out[blockIdx.x * 256 + threadIdx.x] = s;
}

void launch_generate(float *A, int n, long long ld, u64 seed, int kind, int col0, int ncols, cudaStream_t st) {
    dim3 grid(n, (ncols + 255) / 256);
    generate_kernel<<<grid, 256, 0, st>>>(A, n, ld, seed, col0, ncols);
    if (kind == 1) diagdom_kernel<<<(n + 127) / 128, 128, 0, st>>>(A, n, ld, seed, col0, ncols);
}

void launch_generate_batched(float *A, int n, long long first, long long count, u64 seed0, cudaStream_t st) {
    generate_batched_kernel<<<148 * 16, 256, 0, st>>>(A, n, first, count, seed0);
}

template <typename T>
static cudaError_t run_residual_t(const T *A, const T *X, int n, double *out_host, int nout, cudaStream_t st) {
    double *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4 * sizeof(double));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d, 0, 4 * sizeof(double), st);
    dim3 grid((n + 63) / 64, (n + 63) / 64);
    residual_kernel<T><<<grid, 256, 0, st>>>(A, X, n, d);
    e = cudaMemcpyAsync(out_host, d, nout * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    return e;
}

cudaError_t run_residual(const float *A, const float *X, int n, double *out_host, cudaStream_t st) {
    return run_residual_t<float>(A, X, n, out_host, 3, st);
}

cudaError_t run_residual_f64(const double *A, const double *X, int n, double *out_host, cudaStream_t st) {
    return run_residual_t<double>(A, X, n, out_host, 4, st);
}

cudaError_t run_ffma_peak(double *tflops, cudaStream_t st) {
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 148 * 8 * 1024 * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d, 0, 148 * 8 * 1024 * sizeof(float), st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8;
    ffma_peak_kernel<<<grid, 1024, 0, st>>>(d, 64, 1.0f);  // warm-up
    ffma_tile_kernel<<<grid * 4, 256, 0, st>>>(d + 1024, d, 64);
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        const int iters = 4096;
        cudaEventRecord(e0, st);
        if (rep & 1) ffma_tile_kernel<<<grid * 4, 256, 0, st>>>(d + 1024, d, iters);
        else ffma_peak_kernel<<<grid, 1024, 0, st>>>(d, iters, 1.0f);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * 1024.0 * grid;  // both probes: 64 FFMA / thread / iteration
        const double tf = flops / (ms * 1e-3) / 1e12;
        best = tf > best ? tf : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return e;
}
