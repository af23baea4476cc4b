// Prototype of the NEXT epilogue of the tensor-core trailing update, with the math taken away (round-2 experiment; see
// DESIGN.md section 10 "what bounds it" and tools/tile_rmw_probe.cu for the register-carried baseline: 0.49 ms at one
// CTA per SM, 0.345 ms linear, N = 16384).
//
// Question: with ONE CTA per SM (the strip kernel's situation: TMEM and shared memory are spent on operands) and the W
// tiles landed in shared memory by TMA instead of registers, DEPTH tiles deep, does a read-modify-write of W reach the
// streaming rate?  Each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the (n/128)^2 tile grid:
//   warp 0 lane 0   producer: waits for a free slot, then 128 x cp.async.bulk (512 B, one per row) global -> shared,
//                   completion on the slot's mbarrier
//   warps 1-4       consumers: wait for the slot, subtract 1 from its 64 KiB in place (stand-in for W - D), make the
//                   writes visible to the async proxy, then thread 0 of them issues 128 x cp.async.bulk shared -> global,
//                   waits until the stores have READ the slot and releases it
// Reported per DEPTH in {1, 2, 3}: ms per pass over W and TB/s (read + write bytes).
// NOT validated on hardware yet (written after the round-1 GPU budget was spent); it checks its own result (every
// element must have decreased by exactly `passes`) and every wait is bounded, so a mistake reads as FAIL / trap, not a hang.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tile_rmw_tma_probe.cu -o tools/tile_rmw_tma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    const long long t0 = clock64();
    for (unsigned spins = 0;; spins++) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spins & 1023u) == 1023u && clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

constexpr int TILE = 128;
// PADDED = 0: slot rows packed (512 B), consumers walk the slot linearly.
// PADDED = 1: slot rows 528 B apart and every consumer thread owns one ROW (lane = row, like the TMEM accumulator layout):
//             the access pattern of an epilogue that subtracts tcgen05.ld data in place, with no transpose; the 16-byte
//             pad makes the 8 lanes of a quarter-warp hit distinct 16-byte bank groups.
template <int DEPTH, int PADDED>
__global__ void __launch_bounds__(160, 1) tile_rmw_tma(float *w, long long ld, int tiles_x, int ntiles) {
    constexpr unsigned ROW = PADDED ? 528u : 512u;
    constexpr int SLOT_BYTES = TILE * (int)ROW;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_full[DEPTH], bar_empty[DEPTH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned slots = smem_u32(smem);
    if (tid == 0) {
        for (int s = 0; s < DEPTH; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < my_tiles; it++) {
                const int t = blockIdx.x + it * gridDim.x, s = it % DEPTH;
                if (it >= DEPTH) mbar_wait(&bar_empty[s], (unsigned)((it / DEPTH) - 1) & 1u);
                float *tile = w + (long long)(t / tiles_x) * TILE * ld + (long long)(t % tiles_x) * TILE;
                mbar_expect_tx(&bar_full[s], TILE * 512u);
                for (int r = 0; r < TILE; r++)
                    bulk_g2s(slots + (unsigned)s * SLOT_BYTES + (unsigned)r * ROW, tile + (long long)r * ld, 512u, &bar_full[s]);
            }
        }
        __syncwarp();
    } else {
        const int ct = tid - 32;  // 0..127
        for (int it = 0; it < my_tiles; it++) {
            const int t = blockIdx.x + it * gridDim.x, s = it % DEPTH;
            mbar_wait(&bar_full[s], (unsigned)(it / DEPTH) & 1u);
            float4 *slot = reinterpret_cast<float4 *>(smem + (size_t)s * SLOT_BYTES);
            if (PADDED) {
                float4 *row = slot + (size_t)ct * (ROW / 16);  // thread = row
#pragma unroll 8
                for (int c = 0; c < 32; c++) {
                    float4 v = row[c];
                    v.x -= 1.f; v.y -= 1.f; v.z -= 1.f; v.w -= 1.f;
                    row[c] = v;
                }
            } else {
#pragma unroll 8
                for (int e = ct; e < SLOT_BYTES / 16; e += 128) {  // consecutive threads, consecutive 16 B: conflict-free
                    float4 v = slot[e];
                    v.x -= 1.f; v.y -= 1.f; v.z -= 1.f; v.w -= 1.f;
                    slot[e] = v;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the bulk store
            asm volatile("bar.sync 1, 128;" ::: "memory");                // the four consumer warps only
            if (ct == 0) {
                float *tile = w + (long long)(t / tiles_x) * TILE * ld + (long long)(t % tiles_x) * TILE;
                for (int r = 0; r < TILE; r++) bulk_s2g(tile + (long long)r * ld, slots + (unsigned)s * SLOT_BYTES + (unsigned)r * ROW, 512u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slot has been read: it may be refilled
                mbar_arrive(&bar_empty[s]);
            }
        }
        if (ct == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA exits
    }
}

template <int DEPTH, int PADDED>
static void run(float *w, int n, int reps, int *passes) {
    constexpr int SLOT_BYTES = TILE * (PADDED ? 528 : 512);
    const int tiles_x = n / TILE, ntiles = tiles_x * tiles_x;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // one CTA per SM even at DEPTH = 1: pad the request so that two CTAs never fit
    const int smem = DEPTH * SLOT_BYTES > 120 * 1024 ? DEPTH * SLOT_BYTES : 120 * 1024;
    cudaFuncSetAttribute(tile_rmw_tma<DEPTH, PADDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    tile_rmw_tma<DEPTH, PADDED><<<sms, 160, smem>>>(w, n, tiles_x, ntiles);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) tile_rmw_tma<DEPTH, PADDED><<<sms, 160, smem>>>(w, n, tiles_x, ntiles);
    cudaEventRecord(e1);
    const cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *passes += reps + 1;
    const double bytes = 2.0 * n * (double)n * 4;
    printf("n=%d  TMA-staged 128 x 128 tiles, 1 CTA/SM, %d deep (%3d KiB of smem), %s: %.3f ms  %.2f TB/s   %s\n", n, DEPTH,
           DEPTH * SLOT_BYTES / 1024, PADDED ? "padded rows, thread = row" : "packed rows, linear walk  ", ms / reps, bytes / (ms / reps * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384, reps = 20;
    float *w = nullptr;
    if (n % TILE || cudaMalloc(&w, (size_t)n * n * 4) != cudaSuccess) { printf("bad n / alloc failed\n"); return 1; }
    cudaMemset(w, 0, (size_t)n * n * 4);
    int passes = 0;
    run<1, 0>(w, n, reps, &passes);
    run<2, 0>(w, n, reps, &passes);
    run<3, 0>(w, n, reps, &passes);
    run<1, 1>(w, n, reps, &passes);
    run<2, 1>(w, n, reps, &passes);
    run<3, 1>(w, n, reps, &passes);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("FAIL: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    // self-check on a sample of rows: every element was decremented exactly `passes` times
    std::vector<float> row(n);
    long long bad = 0;
    for (int i = 0; i < n; i += 997) {
        cudaMemcpy(row.data(), w + (size_t)i * n, (size_t)n * 4, cudaMemcpyDeviceToHost);
        for (int j = 0; j < n; j++) bad += (row[j] != -(float)passes);
    }
    printf("%s (%lld wrong elements in the sampled rows, expected value %d)\n", bad ? "FAIL" : "PASS", bad, -passes);
    return bad != 0;
}
