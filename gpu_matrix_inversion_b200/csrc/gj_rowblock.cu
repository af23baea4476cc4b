// Row interchanges + row-block recurrence for every column outside the current panel
// (SURVEY.md Appendix A.4 steps 2 and 3).
//
// For a panel with pivot rows k0..k0+kb-1 the reference would, column step by column step,
//   swap rows r <-> p            pivotElementsKernel (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:154-173)
//   divide row r by the pivot    fixRowKernel        (:138-150)
//   eliminate with row r         fixColumnKernel     (:13-57)
// on ALL columns.  Here the kb swaps are applied at once as a gather (the net permutation is kept in
// PanelState by the panel kernels), and the part of the elimination that involves only the kb pivot
// rows is replayed per column in exactly the reference's order:
//   for t: u = x[t] / v_t;  U[t][j] = u (snapshot);  x[t] = u;  x[t2] = fma(-C[t2][t], u, x[t2]) (t2 != t)
// The snapshot U is the B operand of the trailing update; x is written back as the new pivot rows.
//
// One CTA owns CW columns and keeps the 128 x CW pivot-row tile in shared memory.  The 128 steps
// are processed as 8 sub-blocks of 16: (a) a 16-step recurrence per column in registers, (b) a
// rank-16 update of the other pivot rows as a small register-tiled contraction.  Same FMA chains,
// ~25x less time than one barrier pair per step.  The kernel sits on the critical path of every
// panel and a CTA's work is a dependency chain (128 divisions) plus 128*128*CW FMAs on one SM, so CW
// is chosen to spread it: 32 columns per CTA up to n = 8192 (one CTA per SM at n = 4096), 64 above.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

#define RBK_TILE 128 // granularity of the skip range (one panel width)
#define RBK_CPLD 132 // padded leading dimension of cpT (keeps rows 16-byte aligned)

template <int CW>
struct __align__(16) RowblockSmem {
    float x[MATINV_NB][CW];           // the kb pivot rows after the swaps, updated in place
    float cpT[MATINV_NB][RBK_CPLD];   // cpT[t2][t] = multiplier of pivot row t2 at step t
    float us[16][CW];                 // U snapshot of the current sub-block
    float cd[MATINV_NB / 16][16][16]; // diagonal 16x16 blocks of cpT, [step][row]: one step's multipliers are contiguous
    float pv[MATINV_NB];
    int pos[2 * MATINV_NB];
    int content[2 * MATINV_NB];
};

template <int CW>
__global__ void __launch_bounds__(256, CW <= 64 ? 2 : 1)
rowblock_kernel(float *__restrict__ W, long long ld, int k0, int kb, int skip_tile, int skip_n, const float *__restrict__ CmT,
                long long ldc, const float *__restrict__ pvg, const PanelState *__restrict__ ps, float *__restrict__ U,
                long long ldu) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RowblockSmem<CW> &s = *reinterpret_cast<RowblockSmem<CW> *>(smem_raw);
    constexpr int NF = CW / 4;        // float4 groups per row
    constexpr int RPP = 256 / NF;     // rows per pass of the CTA
    constexpr int NP = MATINV_NB / RPP;
    const int j0 = blockIdx.x * CW;
    const int tile = j0 / RBK_TILE;
    if (tile >= skip_tile && tile < skip_tile + skip_n) return;  // panel columns (and look-ahead block)
    const int tid = threadIdx.x, tx = tid % NF, ty = tid / NF;
    const int m = ps->m;

    for (int i = tid; i < 2 * MATINV_NB; i += 256) { s.pos[i] = ps->pos[i]; s.content[i] = ps->content[i]; }
    if (tid < MATINV_NB) s.pv[tid] = (tid < kb) ? pvg[tid] : 1.0f;
    // multipliers of the pivot rows, transposed; steps beyond kb are zeroed (fma(-0, 0, a) == a exactly)
    for (int e = tid; e < MATINV_NB * 32; e += 256) {
        const int t = e & (MATINV_NB - 1), f = e >> 7;  // lanes along t: conflict-free transposed stores
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < kb) v = *reinterpret_cast<const float4 *>(CmT + (long long)t * ldc + k0 + 4 * f);
        s.cpT[4 * f + 0][t] = v.x; s.cpT[4 * f + 1][t] = v.y; s.cpT[4 * f + 2][t] = v.z; s.cpT[4 * f + 3][t] = v.w;
    }
    __syncthreads();
    for (int e = tid; e < MATINV_NB * 16; e += 256) {
        const int j = e & 15, t = (e >> 4) & 15, sbk = e >> 8;
        s.cd[sbk][t][j] = s.cpT[16 * sbk + j][16 * sbk + t];
    }

    // ---- the kb row interchanges as one gather: every read happens before the first write
    float *wc = W + j0 + 4 * tx;
    float4 outside[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const int t = ty + RPP * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < kb) v = *reinterpret_cast<const float4 *>(wc + (long long)s.pos[s.content[t]] * ld);
        *reinterpret_cast<float4 *>(&s.x[t][4 * tx]) = v;
        const int idx = kb + t;
        if (idx < m && s.content[idx] != idx)
            outside[k] = *reinterpret_cast<const float4 *>(wc + (long long)s.pos[s.content[idx]] * ld);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const int idx = kb + ty + RPP * k;
        if (idx < m && s.content[idx] != idx) *reinterpret_cast<float4 *>(wc + (long long)s.pos[idx] * ld) = outside[k];
    }

    // ---- recurrence, 16 steps at a time
    const int nsub = (kb + 15) >> 4;
#pragma unroll 1
    for (int sb = 0; sb < nsub; sb++) {
        const int b0 = sb * 16;
        const int sw = (kb - b0 < 16) ? kb - b0 : 16;
        if (tid < CW) {
            float xx[16];
#pragma unroll
            for (int j = 0; j < 16; j++) xx[j] = s.x[b0 + j][tid];
            if (sw == 16) {
                // Full sub-block: the chain per step is division -> FMA of the next pivot row; pivots are loaded up
                // front and the multipliers of step t+1 are fetched (4 LDS.128) before the division of step t.
                float pvr[16], cn[16];
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    const float4 p4 = *reinterpret_cast<const float4 *>(&s.pv[b0 + 4 * f]);
                    pvr[4 * f] = p4.x; pvr[4 * f + 1] = p4.y; pvr[4 * f + 2] = p4.z; pvr[4 * f + 3] = p4.w;
                    const float4 c4 = *reinterpret_cast<const float4 *>(&s.cd[sb][0][4 * f]);
                    cn[4 * f] = c4.x; cn[4 * f + 1] = c4.y; cn[4 * f + 2] = c4.z; cn[4 * f + 3] = c4.w;
                }
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    float c[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) c[j] = cn[j];
                    if (t + 1 < 16) {
#pragma unroll
                        for (int f = 0; f < 4; f++) {
                            const float4 c4 = *reinterpret_cast<const float4 *>(&s.cd[sb][t + 1][4 * f]);
                            cn[4 * f] = c4.x; cn[4 * f + 1] = c4.y; cn[4 * f + 2] = c4.z; cn[4 * f + 3] = c4.w;
                        }
                    }
                    const float u = xx[t] / pvr[t];
                    s.us[t][tid] = u;
                    U[(long long)(b0 + t) * ldu + j0 + tid] = u;
                    xx[t] = u;
#pragma unroll
                    for (int j = 0; j < 16; j++)
                        if (j != t) xx[j] = gj_elim(xx[j], c[j], u);
                }
            } else {
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    if (t < sw) {
                        const float u = xx[t] / s.pv[b0 + t];
                        s.us[t][tid] = u;
                        U[(long long)(b0 + t) * ldu + j0 + tid] = u;
                        xx[t] = u;
#pragma unroll
                        for (int j = 0; j < 16; j++)
                            if (j != t) xx[j] = gj_elim(xx[j], s.cpT[b0 + j][b0 + t], u);   // rows beyond kb: multipliers are 0
                    } else {
                        s.us[t][tid] = 0.0f;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 16; j++) s.x[b0 + j][tid] = xx[j];
        }
        __syncthreads();
        // rank-16 update of the other pivot rows: thread = 4 columns x rows {ty, ty+RPP, ...}
        float4 u4[16];
#pragma unroll
        for (int t = 0; t < 16; t++) u4[t] = *reinterpret_cast<const float4 *>(&s.us[t][4 * tx]);
#pragma unroll 2
        for (int k = 0; k < NP; k++) {
            const int t2 = ty + RPP * k;
            if (t2 >= kb || (t2 >= b0 && t2 < b0 + 16)) continue;
            float4 acc = *reinterpret_cast<const float4 *>(&s.x[t2][4 * tx]);
            float c[16];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                const float4 cv = *reinterpret_cast<const float4 *>(&s.cpT[t2][b0 + 4 * f]);
                c[4 * f] = cv.x; c[4 * f + 1] = cv.y; c[4 * f + 2] = cv.z; c[4 * f + 3] = cv.w;
            }
#pragma unroll
            for (int t = 0; t < 16; t++) {
                acc.x = gj_elim(acc.x, c[t], u4[t].x);
                acc.y = gj_elim(acc.y, c[t], u4[t].y);
                acc.z = gj_elim(acc.z, c[t], u4[t].z);
                acc.w = gj_elim(acc.w, c[t], u4[t].w);
            }
            *reinterpret_cast<float4 *>(&s.x[t2][4 * tx]) = acc;
        }
        __syncthreads();
    }
#pragma unroll 4
    for (int k = 0; k < NP; k++) {
        const int t = ty + RPP * k;
        if (t < kb) *reinterpret_cast<float4 *>(wc + (long long)(k0 + t) * ld) = *reinterpret_cast<const float4 *>(&s.x[t][4 * tx]);
    }
}

template <int CW>
static void launch_rowblock_cw(float *W, long long ld, int ncols_pad, int k0, int kb, int skip_tile, int skip_n, const float *CmT,
                               long long ldc, const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(rowblock_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowblockSmem<CW>));
    }
    rowblock_kernel<CW><<<ncols_pad / CW, 256, sizeof(RowblockSmem<CW>), st>>>(W, ld, k0, kb, skip_tile, skip_n, CmT, ldc, pv, ps, U, ldu);
}

// MATINV_ROWBLOCK_CW = 32 | 64 | 128 overrides the columns per CTA
static int rowblock_cw(int ncols_pad) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("MATINV_ROWBLOCK_CW");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 32 || forced == 64 || forced == 128) return forced;
    return ncols_pad <= 8192 ? 32 : 64;
}

// W: local column storage (ncols_pad columns, a multiple of 128); tiles of 128 columns [skip_tile, skip_tile + skip_n)
// are left alone (skip_n = 0: none).
void launch_rowblock_ex(float *W, long long ld, int ncols_pad, int k0, int kb, int skip_tile, int skip_n, const float *CmT,
                        long long ldc, const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st) {
    if (ncols_pad <= 0) return;
    switch (rowblock_cw(ncols_pad)) {
    case 32: launch_rowblock_cw<32>(W, ld, ncols_pad, k0, kb, skip_tile, skip_n, CmT, ldc, pv, ps, U, ldu, st); break;
    case 64: launch_rowblock_cw<64>(W, ld, ncols_pad, k0, kb, skip_tile, skip_n, CmT, ldc, pv, ps, U, ldu, st); break;
    default: launch_rowblock_cw<128>(W, ld, ncols_pad, k0, kb, skip_tile, skip_n, CmT, ldc, pv, ps, U, ldu, st); break;
    }
}

void launch_rowblock(float *W, long long ld, int ncols_pad, int k0, int kb, const float *CmT, long long ldc,
                     const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st) {
    launch_rowblock_ex(W, ld, ncols_pad, k0, kb, k0 / RBK_TILE, 1, CmT, ldc, pv, ps, U, ldu, st);
}
