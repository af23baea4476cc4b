// What HBM bandwidth does the ACCESS PATTERN of the trailing update allow, with no math at all?
// W (n x n FP32, row-major, 64 KiB between rows at n = 16384) is read-modified-written once, either linearly or in
// 128-row tiles of TW columns (TW * 4 contiguous bytes per row, rows n*4 bytes apart) -- the pattern of gj_gemm_tc.cu's
// epilogue (4 rows x 128 B per warp instruction, a whole tile in flight per CTA).  If the tiled pattern is far below the
// linear one, the tcgen05 update (0.55 ms = 3.85 TB/s) is bound by DRAM page locality, not by its pipeline.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tile_rmw_probe.cu -o tools/tile_rmw_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(256) linear_rmw(float4 *w, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        float4 v = w[i];
        v.x -= 1.f; v.y -= 1.f; v.z -= 1.f; v.w -= 1.f;
        w[i] = v;
    }
}

// one CTA = 128 rows x TW columns; warp w owns rows 32 (w % 4) .. +31 of column block (w / 4) of 128 columns;
// per instruction 4 rows x 128 B, all 32 loads of the warp's 32 x 128 patch issued before the first store
template <int TW>
__global__ void __launch_bounds__(TW, 1) tile_rmw(float *w, long long ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, cb = warp >> 2;
    float *base = w + ((long long)blockIdx.y * 128 + q * 32 + (lane >> 3)) * ld + (long long)blockIdx.x * TW + cb * 128 + (lane & 7) * 4;
    float4 c[32];
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
#pragma unroll
        for (int i = 0; i < 8; i++) c[ch * 8 + i] = *reinterpret_cast<const float4 *>(base + (long long)(4 * i) * ld + 32 * ch);
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v = c[ch * 8 + i];
            v.x -= 1.f; v.y -= 1.f; v.z -= 1.f; v.w -= 1.f;
            *reinterpret_cast<float4 *>(base + (long long)(4 * i) * ld + 32 * ch) = v;
        }
}

// The same 128 x 128 tile when W is stored PANEL-MAJOR (W[tj][i][jj]: every tile column an n x 128 row-major matrix, so a
// tile is 64 KiB contiguous and the panel kernels keep their (pointer, ld = 128) addressing): what a layout change buys.
__global__ void __launch_bounds__(128, 1) tile_rmw_panel_major(float *w, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *base = w + (long long)blockIdx.x * n * 128 + ((long long)blockIdx.y * 128 + warp * 32 + (lane >> 3)) * 128 + (lane & 7) * 4;
    float4 c[32];
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
#pragma unroll
        for (int i = 0; i < 8; i++) c[ch * 8 + i] = *reinterpret_cast<const float4 *>(base + (4 * i) * 128 + 32 * ch);
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v = c[ch * 8 + i];
            v.x -= 1.f; v.y -= 1.f; v.z -= 1.f; v.w -= 1.f;
            *reinterpret_cast<float4 *>(base + (4 * i) * 128 + 32 * ch) = v;
        }
}

static float time_panel_major(float *w, int n, int reps, int occ) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dim3 grid(n / 128, n / 128);
    const int smem = 220 * 1024 / occ - 2048;
    cudaFuncSetAttribute(tile_rmw_panel_major, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tile_rmw_panel_major<<<grid, 128, smem>>>(w, n);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) tile_rmw_panel_major<<<grid, 128, smem>>>(w, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

// occ = CTAs per SM, enforced through an unused dynamic shared-memory request (the depth of loads in flight per SM is
// occ x TW/128 x 64 KiB; the tcgen05 strip kernel has 64 KiB, the tile kernel 2 x 32 KiB)
template <int TW>
static float time_tile(float *w, int n, int reps, int occ) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dim3 grid(n / TW, n / 128);
    const int smem = 220 * 1024 / occ - 2048;
    cudaFuncSetAttribute(tile_rmw<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    tile_rmw<TW><<<grid, TW, smem>>>(w, n);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) tile_rmw<TW><<<grid, TW, smem>>>(w, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384, reps = 20;
    float *w = nullptr;
    if (cudaMalloc(&w, (size_t)n * n * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(w, 0, (size_t)n * n * 4);
    const double bytes = 2.0 * n * (double)n * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    linear_rmw<<<148 * 8, 256>>>((float4 *)w, (size_t)n * n / 4);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) linear_rmw<<<148 * 8, 256>>>((float4 *)w, (size_t)n * n / 4);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("n=%d  linear RMW            : %.3f ms  %.2f TB/s\n", n, ms / reps, bytes / (ms / reps * 1e-3) / 1e12);
    const int occs[] = {1, 2, 3};
    for (int occ : occs) {
        const float t = time_tile<128>(w, n, reps, occ);
        printf("n=%d  tiles 128 x 128  (512 B per row), %d CTA/SM (%3d KiB in flight): %.3f ms  %.2f TB/s\n", n, occ, occ * 64, t, bytes / (t * 1e-3) / 1e12);
    }
    for (int occ = 1; occ <= 2; occ++) {
        const float t = time_tile<256>(w, n, reps, occ);
        printf("n=%d  tiles 128 x 256  (1 KiB per row), %d CTA/SM (%3d KiB in flight): %.3f ms  %.2f TB/s\n", n, occ, occ * 128, t, bytes / (t * 1e-3) / 1e12);
    }
    {
        const float t = time_tile<512>(w, n, reps, 1);
        printf("n=%d  tiles 128 x 512  (2 KiB per row), 1 CTA/SM (256 KiB in flight): %.3f ms  %.2f TB/s\n", n, t, bytes / (t * 1e-3) / 1e12);
    }
    for (int occ = 1; occ <= 3; occ++) {
        const float t = time_panel_major(w, n, reps, occ);
        printf("n=%d  tiles 128 x 128 of a PANEL-MAJOR W (64 KiB contiguous), %d CTA/SM: %.3f ms  %.2f TB/s\n", n, occ, t, bytes / (t * 1e-3) / 1e12);
    }
    for (int occ : occs) {  // the first rows again, to see how repeatable the numbers are
        const float t = time_tile<128>(w, n, reps, occ);
        printf("n=%d  (again) tiles 128 x 128, %d CTA/SM: %.3f ms  %.2f TB/s\n", n, occ, t, bytes / (t * 1e-3) / 1e12);
    }
    {
        const float t = time_tile<256>(w, n, reps, 2);
        printf("n=%d  (again) tiles 128 x 256, 2 CTA/SM: %.3f ms  %.2f TB/s\n", n, t, bytes / (t * 1e-3) / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
