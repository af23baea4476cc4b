// Micro-benchmark: shared-memory wavefront cost of LDS.128 broadcast patterns on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/lds_probe tools/lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int pat_index(int lane, int p) {
    switch (p) {
        case 0: return 0;
        case 1: return lane / 8;        // 4 distinct, 8 consecutive lanes share      (GEMM A fragment)
        case 2: return lane % 8;        // 8 distinct, lanes l, l+8, l+16, l+24 share (GEMM B fragment)
        case 3: return lane / 4;        // 8 distinct, 4 consecutive lanes share
        case 4: return lane % 4;        // 4 distinct, stride-4 lanes share
        case 5: return lane / 16;       // 2 distinct
        case 6: return lane % 16;       // 16 distinct, 2 share
        case 7: return lane;            // 32 distinct
        case 8: return lane / 2;        // 16 distinct, pairs share
        case 9: return (lane % 8) * 2;  // 8 distinct, 32B apart, stride-8 share
        case 10: return (lane / 8) * 8; // 4 distinct 128B apart
        case 11: return ((lane & 3) | ((lane >> 4) << 2)); // 8 distinct: lanes {l, l+4, l+8, l+12} share
        default: return 0;
    }
}

__global__ void probe(float *out, long long *cyc, int p, int iters) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int idx = pat_index(lane, p);
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const float4 v = sm[(idx + 32 * u + (it & 1) * 512) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc.x == 1234.5f) out[threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8 * 148);
    const int iters = 2000;
    for (int warps : {1, 16}) {
        for (int p = 0; p <= 11; p++) {
            probe<<<1, 32 * warps>>>(out, cyc, p, 10);
            probe<<<1, 32 * warps>>>(out, cyc, p, iters);
            long long c = 0;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps=%2d pattern=%2d cycles/LDS.128(per warp-instr, SM-wide)=%.2f\n", warps, p,
                   (double)c / (iters * 16.0 * warps));
        }
    }
    return 0;
}
