// Row interchanges + row-block recurrence for every column outside the current panel
// (SURVEY.md Appendix A.4 steps 2 and 3).
//
// For a panel with pivot rows k0..k0+kb-1 the reference would, column step by column step,
//   swap rows r <-> p            pivotElementsKernel (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:154-173)
//   divide row r by the pivot    fixRowKernel        (:138-150)
//   eliminate with row r         fixColumnKernel     (:13-57)
// on ALL columns.  Here the kb swaps are applied at once as a gather (the net permutation is kept
// in PanelState by the panel kernel), and the part of the elimination that involves only the kb
// pivot rows is replayed per column in exactly the reference's order:
//   for t: u = x[t] / v_t;  U[t][j] = u (snapshot);  x[t] = u;  x[t2] = fma(-C[t2][t], u, x[t2]) (t2 != t)
// The snapshot U is the B operand of the trailing update, x is written back as the new pivot rows.
#include "common.cuh"
#include "kernels.h"

#define RBK_CW 32  // columns per CTA

struct RowblockSmem {
    float old_[2 * MATINV_NB][RBK_CW];  // original contents of every slot touched by the swaps
    float x[MATINV_NB][RBK_CW];         // the kb pivot rows after the swaps, updated in place
    float cp[MATINV_NB][MATINV_NB];     // cp[t][t2] = multiplier of pivot row t2 at step t
    float pv[MATINV_NB];
    int pos[2 * MATINV_NB];
    int content[2 * MATINV_NB];
};

__global__ void __launch_bounds__(256)
rowblock_kernel(float *__restrict__ W, long long ld, int k0, int kb, const float *__restrict__ CmT, long long ldc,
                const float *__restrict__ pvg, const PanelState *__restrict__ ps, float *__restrict__ U,
                long long ldu) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RowblockSmem &s = *reinterpret_cast<RowblockSmem *>(smem_raw);
    const int j0 = blockIdx.x * RBK_CW;
    if (j0 >= k0 && j0 < k0 + MATINV_NB) return;  // the panel's own columns were handled by the panel kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = ps->m;

    for (int i = threadIdx.x; i < 2 * MATINV_NB; i += 256) { s.pos[i] = ps->pos[i]; s.content[i] = ps->content[i]; }
    if (threadIdx.x < kb) s.pv[threadIdx.x] = pvg[threadIdx.x];
    for (int i = threadIdx.x; i < kb * kb; i += 256) {
        const int t = i / kb, t2 = i - t * kb;
        s.cp[t][t2] = CmT[(long long)t * ldc + k0 + t2];
    }
    __syncthreads();
    for (int idx = warp; idx < m; idx += 8) s.old_[idx][lane] = W[(long long)s.pos[idx] * ld + j0 + lane];
    __syncthreads();
    // rows displaced out of the pivot block
    for (int idx = kb + warp; idx < m; idx += 8) {
        const int c = s.content[idx];
        if (c != idx) W[(long long)s.pos[idx] * ld + j0 + lane] = s.old_[c][lane];
    }
    for (int t = warp; t < kb; t += 8) s.x[t][lane] = s.old_[s.content[t]][lane];
    __syncthreads();

    for (int t = 0; t < kb; t++) {
        const float u = s.x[t][lane] / s.pv[t];
        __syncthreads();  // everyone has read x[t] before its owner overwrites it
        if ((t & 7) == warp) {
            s.x[t][lane] = u;
            U[(long long)t * ldu + j0 + lane] = u;
        }
        for (int t2 = warp; t2 < kb; t2 += 8)
            if (t2 != t) s.x[t2][lane] = gj_elim(s.x[t2][lane], s.cp[t][t2], u);
        __syncthreads();
    }
    for (int t = warp; t < kb; t += 8) W[(long long)(k0 + t) * ld + j0 + lane] = s.x[t][lane];
}

void launch_rowblock(float *W, long long ld, int ncols_pad, int k0, int kb, const float *CmT, long long ldc,
                     const float *pv, const PanelState *ps, float *U, long long ldu, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(rowblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowblockSmem));
        configured = true;
    }
    rowblock_kernel<<<ncols_pad / RBK_CW, 256, sizeof(RowblockSmem), st>>>(W, ld, k0, kb, CmT, ldc, pv, ps, U, ldu);
}
