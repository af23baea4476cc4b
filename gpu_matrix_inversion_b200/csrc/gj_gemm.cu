// Trailing update of the blocked Gauss-Jordan:
//
//     W[i][j] <- fma(-C[i][kb-1], U[kb-1][j], ... fma(-C[i][1], U[1][j], fma(-C[i][0], U[0][j], W[i][j])))
//
// for every row i outside the kb pivot rows and every column j outside the panel.  This is the
// reference's fixColumnKernel (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:13-57)
// applied kb times, with the kb passes over the matrix collapsed into one: a dense FP32
// contraction with K = kb whose accumulator is SEEDED with the C element and walks k in
// order -- the exact FMA chain of the unblocked algorithm, hence bit-identical results and
// bit-identical pivots downstream.  No split-K, no `a - sum`, no tensor-core reordering.
//
// FP32 SIMT GEMM, 128x128 CTA tile, 8x8 register tile per thread, K chunks staged in shared
// memory by a cp.async ring.  Operands are both K-major in HBM (CmT[t][i], U[t][j]) so the
// staging is a straight copy and the inner loop is 4 LDS.128 + 64 FFMA per k.
//
// Lane mapping (MAP = 1): measured on B200 (tools/lds_probe.cu), an LDS.128 whose equal addresses
// sit in CONSECUTIVE lanes costs 2 cycles, the same data shared by lanes 8 apart costs 4.  Each
// quarter-warp is therefore a 2(m) x 4(n) patch of the 4 x 8 lane grid, so both the A and the B
// fragment loads see their duplicates inside a quarter-warp.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define GT 128  // tile edge

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Packed FP32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): two IEEE-rn FMAs per lane in ONE issue slot, so the LDS.128
// fragment loads no longer compete with the FMAs for issue slots (measured with tools/ffma2_probe.cu on B200: 64.9
// TFLOP/s for this inner loop against 56.8 with scalar FFMA; FFMA-only peak 72.5).  Each half is a plain fmaf, hence
// bit-identical results.  ptxas folds the (a,a) pair and the negation into the scalar-broadcast operand: -Ra.F32.
typedef unsigned long long u64p;
__device__ __forceinline__ u64p pack2(float lo, float hi) { u64p r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64p fma2(u64p a, u64p b, u64p c) { u64p d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int BK, int STAGES>
struct GemmSmem {
    float a[STAGES][BK][GT];
    float b[STAGES][BK][GT];
};

template <int BK, int STAGES, int MAP>
__global__ void __launch_bounds__(256, 2)
trailing_gemm_kernel(float *__restrict__ W, long long ld, int row_skip, int col_skip, int col_skip_n, int kb,
                     const float *__restrict__ CmT,
                     long long ldc, const float *__restrict__ U, long long ldu) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GemmSmem<BK, STAGES> &s = *reinterpret_cast<GemmSmem<BK, STAGES> *>(smem_raw);

    int tj = blockIdx.x, ti = blockIdx.y;
    if (col_skip >= 0 && tj >= col_skip) tj += col_skip_n;  // skip the panel's tile column(s) (absent on non-owner shards)
    ti += (ti >= row_skip);  // skip the pivot rows' tile row
    const long long i0 = (long long)ti * GT, j0 = (long long)tj * GT;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    int lm, ln;
    if (MAP == 0) {
        lm = lane >> 3; ln = lane & 7;
    } else {
        const int q = lane >> 3, r8 = lane & 7;
        lm = (q >> 1) * 2 + (r8 >> 2);
        ln = (q & 1) * 4 + (r8 & 3);
    }
    const int rm = wm * 32 + lm * 4;  // thread rows: rm+{0..3}, rm+16+{0..3}
    const int cn = wn * 64 + ln * 4;  // thread cols: cn+{0..3}, cn+32+{0..3}

    // ---- operand staging: BK x 128 floats per operand per chunk, 8 rows per pass
    const int lrow = tid >> 5, lcol = (tid & 31) * 4;
    const float *ga = CmT + i0 + lcol;
    const float *gb = U + j0 + lcol;
    const int nchunks = (kb + BK - 1) / BK;
    auto issue = [&](int kc) {
        const int st = kc % STAGES;
#pragma unroll
        for (int pss = 0; pss < BK / 8; pss++) {
            const int rr = lrow + 8 * pss;
            const long long k = (long long)kc * BK + rr;
            cp_async16(&s.a[st][rr][lcol], ga + k * ldc);
            cp_async16(&s.b[st][rr][lcol], gb + k * ldu);
        }
    };
#pragma unroll
    for (int kc = 0; kc < STAGES - 1; kc++) {
        if (kc < nchunks) issue(kc);
        cp_async_commit();
    }

    // ---- accumulators seeded from C, held as 8 x 4 packed pairs (columns j, j+1)
    u64p acc[8][4];
    float *wp = W + (i0 + rm) * ld + j0 + cn;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float *row = wp + (long long)((i & 3) + (i >> 2) * 16) * ld;
        const ulonglong2 c0 = *reinterpret_cast<const ulonglong2 *>(row);
        const ulonglong2 c1 = *reinterpret_cast<const ulonglong2 *>(row + 32);
        acc[i][0] = c0.x; acc[i][1] = c0.y; acc[i][2] = c1.x; acc[i][3] = c1.y;
    }

#define GEMM_K_STEP(k)                                                                                    \
    {                                                                                                     \
        const float4 a0 = *reinterpret_cast<const float4 *>(&s.a[st][k][rm]);                             \
        const float4 a1 = *reinterpret_cast<const float4 *>(&s.a[st][k][rm + 16]);                        \
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&s.b[st][k][cn]);                     \
        const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(&s.b[st][k][cn + 32]);                \
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};                              \
        const u64p b[4] = {b0.x, b0.y, b1.x, b1.y};                                                       \
        _Pragma("unroll") for (int i = 0; i < 8; i++) {                                                   \
            const u64p na = pack2(-a[i], -a[i]);                                                          \
            _Pragma("unroll") for (int j = 0; j < 4; j++) acc[i][j] = fma2(na, b[j], acc[i][j]);          \
        }                                                                                                 \
    }

    for (int kc = 0; kc < nchunks; kc++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (kc + STAGES - 1 < nchunks) issue(kc + STAGES - 1);
        cp_async_commit();
        const int st = kc % STAGES;
        const int kmax = min(BK, kb - kc * BK);
        if (kmax == BK) {
#pragma unroll
            for (int k = 0; k < BK; k++) GEMM_K_STEP(k)
        } else {
            for (int k = 0; k < kmax; k++) GEMM_K_STEP(k)
        }
    }
    cp_async_wait<0>();
#undef GEMM_K_STEP

#pragma unroll
    for (int i = 0; i < 8; i++) {
        float *row = wp + (long long)((i & 3) + (i >> 2) * 16) * ld;
        *reinterpret_cast<ulonglong2 *>(row) = make_ulonglong2(acc[i][0], acc[i][1]);
        *reinterpret_cast<ulonglong2 *>(row + 32) = make_ulonglong2(acc[i][2], acc[i][3]);
    }
}

template <int BK, int STAGES, int MAP>
static void launch_variant(float *W, long long ld, int nrow_tiles, int ncol_tiles, int row_skip, int col_skip, int col_skip_n, int kb,
                           const float *CmT, long long ldc, const float *U, long long ldu, cudaStream_t st) {
    static bool configured[64] = {};
    const int smem = (int)sizeof(GemmSmem<BK, STAGES>);
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(trailing_gemm_kernel<BK, STAGES, MAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    }
    const int gx = ncol_tiles - (col_skip >= 0 ? col_skip_n : 0), gy = nrow_tiles - 1;
    if (gx <= 0 || gy <= 0) return;
    dim3 grid(gx, gy);
    trailing_gemm_kernel<BK, STAGES, MAP><<<grid, 256, smem, st>>>(W, ld, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu);
}

// MATINV_GEMM selects a variant (tuning aid); the default is the fastest measured on B200.
static int gemm_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MATINV_GEMM");
        v = e ? atoi(e) : 5;  // 5 = BK 32, 3 stages, straight 4x8 lane grid: fastest with the FFMA2 inner loop on B200
    }
    return v;
}

// nrow_tiles x ncol_tiles tiles of 128; tile row `row_skip` (the pivot rows) and `col_skip_n` tile columns from `col_skip` (the panel,
// -1 when the panel lives on another shard) are left untouched.
void launch_trailing_gemm_ex(float *W, long long ld, int nrow_tiles, int ncol_tiles, int row_skip, int col_skip, int col_skip_n, int kb,
                             const float *CmT, long long ldc, const float *U, long long ldu, cudaStream_t st) {
    switch (gemm_variant()) {
        case 0: launch_variant<16, 3, 0>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        case 1: launch_variant<16, 3, 1>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        case 6: launch_variant<32, 2, 1>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        case 3: launch_variant<16, 4, 1>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        case 4: launch_variant<8, 4, 1>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        case 2: launch_variant<32, 3, 1>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
        default: launch_variant<32, 3, 0>(W, ld, nrow_tiles, ncol_tiles, row_skip, col_skip, col_skip_n, kb, CmT, ldc, U, ldu, st); break;
    }
}

void launch_trailing_gemm(float *W, long long ld, int npad, int k0, int kb, const float *CmT, long long ldc,
                          const float *U, long long ldu, cudaStream_t st) {
    const int nt = npad / GT;
    launch_trailing_gemm_ex(W, ld, nt, nt, k0 / GT, k0 / GT, 1, kb, CmT, ldc, U, ldu, st);
}
