// Optional trailing update on the 5th-generation tensor cores: 3xTF32 with tcgen05.mma and a TMEM accumulator
// (north_star kernel (3), "optional 3xTF32 tcgen05/TMEM variant gated by residual"; MATINV_FLAG_TF32X3).
//
// Same contraction as gj_gemm.cu -- the reference's fixColumnKernel
// (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:13-57) applied 128 times,
//
//     W[i][j] <- W[i][j] - sum_t C[i][t] * U[t][j]        (i outside the pivot rows, j outside the panel)
//
// but NOT the reference's FMA chain: every FP32 operand is split into two TF32 numbers, x = hi + lo
// (hi = rna_tf32(x), lo = rna_tf32(x - hi); 11 + 11 significant bits), and the product is evaluated as
// lo*hi + hi*lo + hi*hi on the tensor cores with an FP32 accumulator that starts from zero.  The dropped lo*lo term and
// the tensor core's accumulation order make the result differ from the FP32 SIMT kernel in the last 2-3 bits, so this
// path can change pivot choices downstream.  It is therefore never the parity path: the shim accepts its result only
// if the residual estimate passes (matinv_shim.cu: tf32x3 gate) and reruns the FP32 SIMT schedule otherwise.
//
// Two launches per trailing update:
//   tf32_split_kernel             reads the K-major operands CmT[t][i] / U[t][j] once and writes, per 128-row tile, the hi and
//                                 lo parts in the exact shared-memory image tcgen05.mma wants (K-major, no swizzle: 8 x 16-byte
//                                 core matrices, 128 contiguous bytes each), so the GEMM kernel stages an operand with linear
//                                 bulk copies and needs no tensor map.  2 x 16 MiB at N=16384, L2-resident.
//   trailing_tf32x3_strip_kernel  one CTA per SM, 192 threads, a strip of column tiles of one tile row per CTA (see below):
//                                   warp 0     allocates 256 TMEM columns; lane 0 issues the MMAs (6 per 16-deep K stage)
//                                   warp 1     lane 0 = producer: cp.async.bulk (TMA, linear) -- the A image once, B through a
//                                              4-slot ring; completion on mbarriers, slots released by tcgen05.commit
//                                   warps 2-5  epilogue: tcgen05.ld -> shared-memory transpose -> W - D in 128-byte row segments
//
// What bounds it: 2 x 64 KiB of W traffic per tile against 3 x 2 x 128^3 tensor flops -- a streaming pass over the 2.1 GB
// of W touched by one trailing update at N=16384 would take 0.33 ms at the measured 6.4 TB/s, the 3xTF32 MMAs 0.18 ms at the
// nominal TF32 peak; the FP32 SIMT kernel is FP32-pipe-bound at 1.19 ms.  Measured: 0.55 ms = 3.85 TB/s of W traffic.  The
// gap to 0.33 ms is the structure of the W traffic, not the MMA pipeline: a pure read-modify-write of W in 128 x 128 tiles
// carried in registers tops out at 0.49 ms at one CTA per SM (tools/tile_rmw_probe.cu; 0.345 ms linear; a panel-major W changes
// little; 128 x 256 tiles at two CTAs per SM, 256 KiB in flight and out of phase, reach 0.322 ms), and neither a deeper operand
// pipeline (the tile kernel below) nor prefetching W a tile ahead moved the number.
// History on B200: one tile per CTA with both operands re-read per tile and a row-per-lane epilogue 0.92 ms; resident A +
// double-buffered accumulator 0.76 ms; coalesced epilogue 0.55 ms.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TC_TILE = 128;                           // tile edge == panel width == MMA M == MMA N
constexpr int TC_BK = 16;                              // K per pipeline stage (2 MMA k-steps of 8)
constexpr int TC_KCH = 128 / TC_BK;                    // stages per tile
constexpr int TC_CHUNK_FLOATS = TC_TILE * TC_BK;       // one operand part (hi or lo) of one stage: 2048 floats
constexpr int TC_CHUNK_BYTES = TC_CHUNK_FLOATS * 4;    // 8 KiB
constexpr int TC_TILE_IMG_FLOATS = TC_KCH * 2 * TC_CHUNK_FLOATS;  // image of one 128-row tile, all K: 128 KiB

// Inside one 8 KiB chunk (128 rows x 16 k, K-major, no swizzle): core matrix = 8 rows x 16 bytes (4 TF32), stored as 128
// contiguous bytes; the 16 row groups of one 4-wide K slice follow each other (SBO = 128 B), the 4 K slices are 2 KiB
// apart (LBO = 2048 B).  Element (r, k) sits at float offset (k/4)*512 + (r/8)*32 + (r%8)*4 + (k%4).
constexpr unsigned TC_LBO = 2048, TC_SBO = 128;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------------------------
// split + re-tile: S is K-major (S[k][c], leading dimension lds); tile t covers columns t*128 .. t*128+127.
// Rows k >= kb are written as zeros (a narrower last panel contributes nothing).
__global__ void __launch_bounds__(256)
tf32_split_kernel(const float *__restrict__ SA, long long ldsa, float *__restrict__ imgA, int tilesA,
                  const float *__restrict__ SB, long long ldsb, float *__restrict__ imgB, int kb) {
    int t = blockIdx.x;
    const float *S = SA;
    long long lds = ldsa;
    float *img = imgA;
    if (t >= tilesA) { t -= tilesA; S = SB; lds = ldsb; img = imgB; }
    const int r = threadIdx.x & 127;
    const float *src = S + (long long)t * TC_TILE + r;
    float *dst = img + (long long)t * TC_TILE_IMG_FLOATS;
#pragma unroll 4
    for (int k4 = threadIdx.x >> 7; k4 < 32; k4 += 2) {
        const int k = 4 * k4;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float x = (k + e < kb) ? src[(long long)(k + e) * lds] : 0.0f;
            hi[e] = to_tf32(x);
            lo[e] = to_tf32(x - hi[e]);
        }
        const int off = ((k >> 4) * 2) * TC_CHUNK_FLOATS + ((k & 15) >> 2) * 512 + (r >> 3) * 32 + (r & 7) * 4;
        *reinterpret_cast<float4 *>(dst + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(dst + off + TC_CHUNK_FLOATS) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 wrappers (inline PTX, sm_100a)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must fail loudly (trap -> CUDA error -> MATINV_E_CUDA), never hang the GPU.
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
        // ~20 s at 2 GHz: far beyond any legitimate wait (a whole inversion takes 0.1-1 s) yet bounded; long enough not to fire
        // under compute-sanitizer, a debugger or time-slicing with another process
        if ((spins & 1023u) == 1023u && clock64() - t0 > 40000000000ll) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], M = N = 128, K = 8, TF32 inputs, FP32 accumulator
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b, unsigned idesc,
                                            unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor (K-major, SWIZZLE_NONE): start address, leading (K) and stride (8-row group) byte
// offsets, all >> 4; bits 46-47 = 1 (sm_100 descriptor version).
__device__ __forceinline__ unsigned long long tc_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr & 0x3FFFFu) >> 4) | ((unsigned long long)(TC_LBO >> 4) << 16) |
           ((unsigned long long)(TC_SBO >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major (bits 15, 16 = 0),
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr unsigned TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TC_TILE >> 3) << 17) | ((unsigned)(TC_TILE >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&d)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
          "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]),
          "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]),
          "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// One CTA per SM walks a STRIP of up to `strip` consecutive column tiles of one tile row.
//   * the A image of the row tile (hi|lo, all K: 128 KiB) is loaded once and stays in shared memory -- re-reading it
//     for every tile doubles the L2 -> SM operand traffic (256 KiB per 128 KiB of W traffic);
//   * only B streams, through a 4-stage ring of 16 KiB stages (K = 16, hi|lo);
//   * two TMEM accumulators (2 x 128 columns): the MMAs of tile t+1 run while the epilogue drains tile t;
//   * the epilogue issues ALL of its W loads for a tile (512 B per thread, 64 KiB per CTA) before it waits for the
//     accumulator, so a full tile of HBM reads is in flight per SM while the tensor core works, and transposes the
//     accumulator through shared memory so that W is read and written in 128-byte row segments.
constexpr int TC2_STAGE_BYTES = 2 * TC_CHUNK_BYTES;                  // B hi | lo for 16 k: 16 KiB
constexpr int TC2_A_BYTES = TC_TILE_IMG_FLOATS * 4;                  // 128 KiB
constexpr unsigned TC2_STAGE_ROW_BYTES = 32 * 4 + 16;               // epilogue patch row: 32 floats + 16 B pad (conflict-free v4 access)
constexpr int TC2_PATCH_BYTES = 32 * (int)TC2_STAGE_ROW_BYTES;       // per epilogue warp: 32 rows
// GROUPS epilogue warp-groups of 4 warps (group g drains accumulator g, i.e. tiles t = g mod 2), STAGES ring slots for B
constexpr int tc2_smem_bytes(int groups, int stages) { return TC2_A_BYTES + stages * TC2_STAGE_BYTES + groups * 4 * TC2_PATCH_BYTES + 1024; }
constexpr int TC2_TMEM_COLS = 256;

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int GROUPS, int TC2_STAGES>
__global__ void __launch_bounds__(64 + 128 * GROUPS, 1)
trailing_tf32x3_strip_kernel(float *__restrict__ W, long long ld, int row_skip, int col_skip, int col_skip_n, int ncols, int strip,
                             const float *__restrict__ imgA, const float *__restrict__ imgB) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    __shared__ __align__(8) unsigned long long bar_a[TC_KCH], bar_full[TC2_STAGES], bar_empty[TC2_STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ unsigned tmem_slot;

    int ti = blockIdx.y;
    ti += (ti >= row_skip);
    const int x0 = blockIdx.x * strip;                         // logical (skip-free) column tile range of this CTA
    const int ntiles = min(strip, ncols - x0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned sA = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const unsigned ring = sA + TC2_A_BYTES;
    const unsigned sStage = ring + TC2_STAGES * TC2_STAGE_BYTES;

    if (tid == 0) {
#pragma unroll
        for (int kc = 0; kc < TC_KCH; kc++) mbar_init(&bar_a[kc], 1);
#pragma unroll
        for (int s = 0; s < TC2_STAGES; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&bar_acc_full[b], 1);
            mbar_init(&bar_acc_empty[b], 128);  // every epilogue thread arrives
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"((unsigned)TC2_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<volatile unsigned *>(&tmem_slot);

    if (warp == 1) {
        // ===== producer
        if (lane == 0) {
            const float *ga = imgA + (long long)ti * TC_TILE_IMG_FLOATS;
            int it = 0;
            for (int t = 0; t < ntiles; t++) {
                int tj = x0 + t;
                if (col_skip >= 0 && tj >= col_skip) tj += col_skip_n;
                const float *gb = imgB + (long long)tj * TC_TILE_IMG_FLOATS;
                for (int kc = 0; kc < TC_KCH; kc++, it++) {
                    const int s = it % TC2_STAGES;
                    if (t == 0) {  // the resident A image travels chunk by chunk next to the first tile's B stages, one
                                   // barrier per chunk, so the first MMAs start after 32 KiB instead of 144 KiB
                        mbar_expect_tx(&bar_a[kc], TC2_STAGE_BYTES);
                        bulk_g2s(sA + (unsigned)kc * TC2_STAGE_BYTES, ga + (long long)kc * 2 * TC_CHUNK_FLOATS, TC2_STAGE_BYTES, &bar_a[kc]);
                    }
                    if (it >= TC2_STAGES) mbar_wait(&bar_empty[s], (unsigned)((it / TC2_STAGES) - 1) & 1u);
                    mbar_expect_tx(&bar_full[s], TC2_STAGE_BYTES);
                    bulk_g2s(ring + (unsigned)s * TC2_STAGE_BYTES, gb + (long long)kc * 2 * TC_CHUNK_FLOATS, TC2_STAGE_BYTES, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        // ===== MMA issuer
        if (lane == 0) {
            int it = 0;
            for (int t = 0; t < ntiles; t++) {
                const int b = t & 1;
                if (t >= 2) mbar_wait(&bar_acc_empty[b], (unsigned)((t >> 1) - 1) & 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const unsigned acc = tmem + (unsigned)b * TC_TILE;
                for (int kc = 0; kc < TC_KCH; kc++, it++) {
                    const int s = it % TC2_STAGES;
                    if (t == 0) mbar_wait(&bar_a[kc], 0);
                    mbar_wait(&bar_full[s], (unsigned)(it / TC2_STAGES) & 1u);
                    tc_fence_after();
                    const unsigned abase = sA + (unsigned)kc * TC2_STAGE_BYTES, bbase = ring + (unsigned)s * TC2_STAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ks++) {
                        const unsigned o = (unsigned)ks * 2u * TC_LBO;
                        const unsigned long long a_hi = tc_desc(abase + o), a_lo = tc_desc(abase + TC_CHUNK_BYTES + o);
                        const unsigned long long b_hi = tc_desc(bbase + o), b_lo = tc_desc(bbase + TC_CHUNK_BYTES + o);
                        tc_mma_tf32(acc, a_lo, b_hi, TC_IDESC, (kc | ks) != 0);
                        tc_mma_tf32(acc, a_hi, b_lo, TC_IDESC, 1u);
                        tc_mma_tf32(acc, a_hi, b_hi, TC_IDESC, 1u);
                    }
                    tc_commit(&bar_empty[s]);
                }
                tc_commit(&bar_acc_full[b]);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue.  TMEM hands every lane one ROW of the accumulator; reading and writing W that way would touch 32
        // different 128-byte lines per instruction.  Each warp therefore transposes its 32 rows through a private, padded
        // shared-memory patch (32 columns at a time) and accesses W as 4 rows x 128 contiguous bytes per instruction.
        const int q = warp & 3, ew = warp - 2, grp = ew >> 2;
        const int rsub = lane >> 3, csub = (lane & 7) * 4;
        const unsigned stage = sStage + (unsigned)ew * TC2_PATCH_BYTES;
        float *wbase = W + ((long long)ti * TC_TILE + q * 32 + rsub) * ld + csub;
        for (int t = grp; t < ntiles; t += GROUPS) {
            const int b = t & 1;
            int tj = x0 + t;
            if (col_skip >= 0 && tj >= col_skip) tj += col_skip_n;
            float *wt = wbase + (long long)tj * TC_TILE;
            float4 c[32];
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
#pragma unroll
                for (int i = 0; i < 8; i++)  // rows 4i + rsub of this warp's 32, columns 32 ch + csub .. +3: all in flight at once
                    c[ch * 8 + i] = *reinterpret_cast<const float4 *>(wt + (long long)(4 * i) * ld + 32 * ch);
            mbar_wait(&bar_acc_full[b], (unsigned)(t >> 1) & 1u);
            tc_fence_after();
            const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)b * TC_TILE;
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                unsigned d[32];
                tmem_ld32(taddr + (unsigned)(32 * ch), d);
                tmem_ld_wait();
#pragma unroll
                for (int v = 0; v < 8; v++)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (unsigned)lane * TC2_STAGE_ROW_BYTES + 16u * v),
                                 "r"(d[4 * v]), "r"(d[4 * v + 1]), "r"(d[4 * v + 2]), "r"(d[4 * v + 3])
                                 : "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float4 dd;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(dd.x), "=f"(dd.y), "=f"(dd.z), "=f"(dd.w)
                                 : "r"(stage + (unsigned)(4 * i + rsub) * TC2_STAGE_ROW_BYTES + 4u * csub)
                                 : "memory");
                    const float4 cc = c[ch * 8 + i];
                    *reinterpret_cast<float4 *>(wt + (long long)(4 * i) * ld + 32 * ch) =
                        make_float4(cc.x - dd.x, cc.y - dd.y, cc.z - dd.z, cc.w - dd.w);
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[b]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)TC2_TMEM_COLS) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL alternative (MATINV_TC_KERNEL=tile; NOT the default, see DESIGN section 10 "what bounds it"): one 128x128 tile
// per CTA, TWO CTAs per SM, both operands streamed through a deeper ring of small stages.  The strip kernel keeps A
// resident, which leaves shared memory for only 64 KiB of B in flight per SM; here nothing is resident, a stage is 8 k
// of all four operand parts (4 x 4 KiB), five slots per CTA -> 160 KiB in flight per SM, and the second CTA's main loop
// overlaps the first one's epilogue instead of a second accumulator.  CTAs live for one tile, which also lets the
// look-ahead panel chain find free SMs sooner.  Costs: A is re-read from L2 for every tile (2 x the operand traffic).
constexpr int TC3_STAGES = 5;
constexpr int TC3_BK = 8;                                  // one MMA k-step per stage
constexpr int TC3_PART_BYTES = TC_TILE * TC3_BK * 4;       // one operand part: 128 rows x 8 k = 4 KiB (two 4-wide K slices)
constexpr int TC3_STAGE_BYTES = 4 * TC3_PART_BYTES;        // A_hi | A_lo | B_hi | B_lo = 16 KiB
constexpr int TC3_SMEM_BYTES = TC3_STAGES * TC3_STAGE_BYTES + 4 * TC2_PATCH_BYTES + 1024;
constexpr int TC3_TMEM_COLS = 128;

__global__ void __launch_bounds__(192, 2)
trailing_tf32x3_tile_kernel(float *__restrict__ W, long long ld, int row_skip, int col_skip, int col_skip_n,
                            const float *__restrict__ imgA, const float *__restrict__ imgB) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    __shared__ __align__(8) unsigned long long bar_full[TC3_STAGES], bar_empty[TC3_STAGES], bar_done;
    __shared__ unsigned tmem_slot;

    int tj = blockIdx.x, ti = blockIdx.y;
    if (col_skip >= 0 && tj >= col_skip) tj += col_skip_n;
    ti += (ti >= row_skip);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned ring = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    const unsigned sStage = ring + TC3_STAGES * TC3_STAGE_BYTES;
    constexpr int NSTEP = 128 / TC3_BK;  // 16 stages per tile

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TC3_STAGES; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"((unsigned)TC3_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *reinterpret_cast<volatile unsigned *>(&tmem_slot);

    if (warp == 1) {
        if (lane == 0) {
            // image of a tile: [k/16][hi|lo][2048 floats]; the two K slices of step kq sit at (kq & 1) * 1024 floats of their chunk
            const float *ga = imgA + (long long)ti * TC_TILE_IMG_FLOATS;
            const float *gb = imgB + (long long)tj * TC_TILE_IMG_FLOATS;
            for (int kq = 0; kq < NSTEP; kq++) {
                const int s = kq % TC3_STAGES;
                if (kq >= TC3_STAGES) mbar_wait(&bar_empty[s], (unsigned)((kq / TC3_STAGES) - 1) & 1u);
                mbar_expect_tx(&bar_full[s], TC3_STAGE_BYTES);
                const unsigned dst = ring + (unsigned)s * TC3_STAGE_BYTES;
                const long long hi = (long long)(kq >> 1) * 2 * TC_CHUNK_FLOATS + (kq & 1) * 1024, lo = hi + TC_CHUNK_FLOATS;
                bulk_g2s(dst, ga + hi, TC3_PART_BYTES, &bar_full[s]);
                bulk_g2s(dst + TC3_PART_BYTES, ga + lo, TC3_PART_BYTES, &bar_full[s]);
                bulk_g2s(dst + 2 * TC3_PART_BYTES, gb + hi, TC3_PART_BYTES, &bar_full[s]);
                bulk_g2s(dst + 3 * TC3_PART_BYTES, gb + lo, TC3_PART_BYTES, &bar_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        if (lane == 0) {
            for (int kq = 0; kq < NSTEP; kq++) {
                const int s = kq % TC3_STAGES;
                mbar_wait(&bar_full[s], (unsigned)(kq / TC3_STAGES) & 1u);
                tc_fence_after();
                const unsigned base = ring + (unsigned)s * TC3_STAGE_BYTES;
                const unsigned long long a_hi = tc_desc(base), a_lo = tc_desc(base + TC3_PART_BYTES);
                const unsigned long long b_hi = tc_desc(base + 2 * TC3_PART_BYTES), b_lo = tc_desc(base + 3 * TC3_PART_BYTES);
                tc_mma_tf32(tmem, a_lo, b_hi, TC_IDESC, kq != 0);
                tc_mma_tf32(tmem, a_hi, b_lo, TC_IDESC, 1u);
                tc_mma_tf32(tmem, a_hi, b_hi, TC_IDESC, 1u);
                tc_commit(&bar_empty[s]);
            }
            tc_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // epilogue as in the strip kernel (TMEM -> padded patch -> 4 rows x 128 B per instruction); the register budget at
        // two CTAs per SM holds half a tile of W, so chunks 2 and 3 are fetched into the registers of chunks 0 and 1
        const int q = warp & 3, ew = warp - 2;
        const int rsub = lane >> 3, csub = (lane & 7) * 4;
        const unsigned stage = sStage + (unsigned)ew * TC2_PATCH_BYTES;
        float *wt = W + ((long long)ti * TC_TILE + q * 32 + rsub) * ld + csub + (long long)tj * TC_TILE;
        float4 c[16];
#pragma unroll
        for (int ch = 0; ch < 2; ch++)
#pragma unroll
            for (int i = 0; i < 8; i++) c[ch * 8 + i] = *reinterpret_cast<const float4 *>(wt + (long long)(4 * i) * ld + 32 * ch);
        mbar_wait(&bar_done, 0);
        tc_fence_after();
        const unsigned taddr = tmem + ((unsigned)(q * 32) << 16);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            unsigned d[32];
            tmem_ld32(taddr + (unsigned)(32 * ch), d);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 8; v++)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (unsigned)lane * TC2_STAGE_ROW_BYTES + 16u * v),
                             "r"(d[4 * v]), "r"(d[4 * v + 1]), "r"(d[4 * v + 2]), "r"(d[4 * v + 3])
                             : "memory");
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float4 dd;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(dd.x), "=f"(dd.y), "=f"(dd.z), "=f"(dd.w)
                             : "r"(stage + (unsigned)(4 * i + rsub) * TC2_STAGE_ROW_BYTES + 4u * csub)
                             : "memory");
                const float4 cc = c[(ch & 1) * 8 + i];
                *reinterpret_cast<float4 *>(wt + (long long)(4 * i) * ld + 32 * ch) =
                    make_float4(cc.x - dd.x, cc.y - dd.y, cc.z - dd.z, cc.w - dd.w);
            }
            if (ch < 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) c[ch * 8 + i] = *reinterpret_cast<const float4 *>(wt + (long long)(4 * i) * ld + 32 * (ch + 2));
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((unsigned)TC3_TMEM_COLS) : "memory");
    }
}

}  // namespace

size_t tf32x3_image_bytes(int tiles) { return (size_t)tiles * TC_TILE_IMG_FLOATS * sizeof(float); }

// Tiles per strip.  Long strips amortise the once-per-CTA cost (TMEM allocation, the 128 KiB A load) but every kernel of the
// look-ahead panel chain on the high-priority stream has to wait until running CTAs retire, so short-lived CTAs matter while
// that chain is the critical path.  Measured on B200 (profiles/README.md): N=16384 strip 1/2/3/4/5/8/12 -> 124.7/107.9/104.8/
// 106.7/108.9/112.3/114.3 ms; N=32768 strip 2/4/8 -> 782/733/706 ms (there the update dominates).  Above 128 tile rows the
// length minimises waves x (L + 1).  MATINV_TC_STRIP overrides.
static int tc_strip_len(int gx, int gy) {
    static int forced = -2;
    if (forced == -2) {
        const char *e = getenv("MATINV_TC_STRIP");
        forced = e ? atoi(e) : -1;
    }
    if (forced > 0) return forced < gx ? forced : gx;
    if (gy <= 128) return gx < 3 ? gx : 3;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int best = 1;
    long long best_cost = -1;
    for (int L = 1; L <= 12 && L <= gx; L++) {
        const long long ctas = (long long)gy * ((gx + L - 1) / L);
        const long long cost = ((ctas + sms - 1) / sms) * (L + 1);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = L; }
    }
    return best;
}

// MATINV_TC_KERNEL=tile selects the experimental one-tile-per-CTA kernel (default: the strip kernel)
static bool tc_tile_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MATINV_TC_KERNEL");
        v = (e && e[0] == 't') ? 1 : 0;
    }
    return v == 1;
}

// Same tile selection as launch_trailing_gemm_ex (gj_gemm.cu).  imgA must hold nrow_tiles tile images, imgB ncol_tiles; both are
// scratch owned by the caller and must not be shared with a launch that may run concurrently on another stream.
cudaError_t launch_trailing_tf32x3(float *W, long long ld, int nrow_tiles, int ncol_tiles, int row_skip, int col_skip, int col_skip_n,
                                   int kb, const float *CmT, long long ldc, const float *U, long long ldu, float *imgA, float *imgB,
                                   cudaStream_t st) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(trailing_tf32x3_strip_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tc2_smem_bytes(1, 4));
        if (e != cudaSuccess) return e;
    }
    const int gx = ncol_tiles - (col_skip >= 0 ? col_skip_n : 0), gy = nrow_tiles - 1;
    if (gx <= 0 || gy <= 0) return cudaSuccess;
    tf32_split_kernel<<<nrow_tiles + ncol_tiles, 256, 0, st>>>(CmT, ldc, imgA, nrow_tiles, U, ldu, imgB, kb);
    if (tc_tile_variant()) {  // experimental, off by default
        static bool configured_tile[64] = {};
        if (first_use_on_device(configured_tile)) {
            cudaError_t e = cudaFuncSetAttribute(trailing_tf32x3_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES);
            if (e != cudaSuccess) return e;
        }
        trailing_tf32x3_tile_kernel<<<dim3(gx, gy), 192, TC3_SMEM_BYTES, st>>>(W, ld, row_skip, col_skip, col_skip_n, imgA, imgB);
        return cudaGetLastError();
    }
    const int L = tc_strip_len(gx, gy);
    // <2, 3> (a second epilogue warp-group, one per accumulator) was measured slower: 0.607 vs 0.550 ms per update at N=16384
    // (320 threads cap the epilogue at 168 registers and spill)
    trailing_tf32x3_strip_kernel<1, 4><<<dim3((gx + L - 1) / L, gy), 192, tc2_smem_bytes(1, 4), st>>>(W, ld, row_skip, col_skip,
                                                                                                     col_skip_n, gx, L, imgA, imgB);
    return cudaGetLastError();
}
