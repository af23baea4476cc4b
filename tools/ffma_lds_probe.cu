// Micro-benchmark: how much of the FFMA peak survives the LDS.128 fragment loads of a register-tiled
// SIMT GEMM inner loop (no global traffic, no barriers).  TM x TN register tile, 256 or 128 threads.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ffma_lds_probe tools/ffma_lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int TM, int TN, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) probe(float *out, int iters) {
    __shared__ __align__(16) float sa[16][128];
    __shared__ __align__(16) float sb[16][128];
    for (int i = threadIdx.x; i < 16 * 128; i += THREADS) { (&sa[0][0])[i] = 1e-3f * i; (&sb[0][0])[i] = 1e-4f * i; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lm = lane >> 3, ln = lane & 7;
    // split fragments (conflict-free): row chunk f at rm + 16 f, column chunk f at cn + 32 f
    const int rm = (warp * 64 + lm * 4) % 128, cn = ln * 4;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k0_ = 0; k0_ < 16; k0_++) {
            const int k = (k0_ + it) & 15;
            float a[TM], b[TN];
#pragma unroll
            for (int f = 0; f < TM / 4; f++) {
                const float4 v = *reinterpret_cast<const float4 *>(&sa[k][(rm + 16 * f) % 128]);
                a[4 * f] = v.x; a[4 * f + 1] = v.y; a[4 * f + 2] = v.z; a[4 * f + 3] = v.w;
            }
#pragma unroll
            for (int f = 0; f < TN / 4; f++) {
                const float4 v = *reinterpret_cast<const float4 *>(&sb[k][(cn + 32 * f) % 128]);
                b[4 * f] = v.x; b[4 * f + 1] = v.y; b[4 * f + 2] = v.z; b[4 * f + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(-a[i], b[j], acc[i][j]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) s += acc[i][j];
    if (s == 12345.678f) out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <int TM, int TN, int THREADS, int MINB>
void run(const char *name, float *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, grid = 148 * MINB * 4;
    probe<TM, TN, THREADS, MINB><<<grid, THREADS>>>(d, 10);
    cudaEventRecord(e0);
    probe<TM, TN, THREADS, MINB><<<grid, THREADS>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * TM * TN * 16.0 * iters * THREADS * grid;
    printf("%-34s %6.2f TFLOP/s  (%s)\n", name, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 16 * 1024 * sizeof(float));
    run<8, 8, 256, 2>("8x8 tile, 256 thr, 2 CTA/SM", d);
    run<8, 8, 256, 1>("8x8 tile, 256 thr, 1 CTA/SM", d);
    run<8, 16, 256, 1>("8x16 tile, 256 thr, 1 CTA/SM", d);
    run<16, 8, 256, 1>("16x8 tile, 256 thr, 1 CTA/SM", d);
    run<8, 16, 128, 2>("8x16 tile, 128 thr, 2 CTA/SM", d);
    run<8, 12, 256, 1>("8x12 tile, 256 thr, 1 CTA/SM", d);
    run<4, 8, 256, 4>("4x8 tile, 256 thr, 4 CTA/SM", d);
    run<8, 8, 128, 4>("8x8 tile, 128 thr, 4 CTA/SM", d);
    return 0;
}
