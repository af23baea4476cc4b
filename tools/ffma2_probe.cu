// Micro-benchmark: packed FP32 FMA (fma.rn.f32x2, sm_100+) in the register-tiled GEMM inner loop.
// 8x8 tile per thread as 8x4 f32x2 accumulators; A values duplicated into (a,a) pairs; B pairs straight from LDS.128.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) probe(float *out, int iters) {
    __shared__ __align__(16) float sa[16][128];
    __shared__ __align__(16) float sb[16][128];
    for (int i = threadIdx.x; i < 16 * 128; i += 256) { (&sa[0][0])[i] = 1e-3f * i; (&sb[0][0])[i] = 1e-4f * i; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lm = lane >> 3, ln = lane & 7;
    const int rm = (warp * 64 + lm * 4) % 128, cn = ln * 4;
    u64 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = pack2((float)(i + j), (float)(i - j));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k0_ = 0; k0_ < 16; k0_++) {
            const int k = (k0_ + it) & 15;
            const float4 a0 = *reinterpret_cast<const float4 *>(&sa[k][rm]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&sa[k][(rm + 16) % 128]);
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&sb[k][cn]);
            const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(&sb[k][(cn + 32) % 128]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const u64 b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const u64 ad = pack2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma2(ad, b[j], acc[i][j]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { float lo, hi; unpack2(acc[i][j], lo, hi); s += lo + hi; }
    if (s == 12345.678f) out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE, int MINB>
void run(const char *name, float *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, grid = 148 * MINB * 4;
    probe<MODE, MINB><<<grid, 256>>>(d, 10);
    cudaEventRecord(e0);
    probe<MODE, MINB><<<grid, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64 * 16.0 * iters * 256 * grid;
    printf("%-34s %6.2f TFLOP/s  (%s)\n", name, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *d;
    cudaMalloc(&d, 148 * 16 * 1024 * sizeof(float));
    run<0, 2>("f32x2 8x8 tile, 256 thr, 2 CTA/SM", d);
    run<0, 1>("f32x2 8x8 tile, 256 thr, 1 CTA/SM", d);
    return 0;
}
