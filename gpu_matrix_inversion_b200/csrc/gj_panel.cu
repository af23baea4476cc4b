// Panel factorisation: in-place Gauss-Jordan restricted to the kb (<=128) columns of one panel,
// all n rows.
//
// One launch per column.  Each launch fuses, for column t of the panel (global column r = k0+t):
//   - the final reduction of the pivot search (partials written by the previous launch)
//     -> replaces finalMaxPivotKernel (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:112-132)
//   - row swap r <-> p                                   -> pivotElementsKernel (:154-173)
//   - pivot-row normalisation by true division            -> fixRowKernel (:138-150)
//   - the rank-1 elimination of the panel columns         -> fixColumnKernel (:13-57)
//   - the partial pivot search of column t+1 on the freshly updated values -> maxPivotKernel (:61-106)
// and records the multipliers C[t][i] (transposed, CmT) for the trailing update.
//
// The panel ping-pongs between two buffers exactly like the reference's two [A|I] buffers
// (:352-359): every element of `out` is a pure function of `in`, so no grid-wide sync is needed.
// Step 0 reads the panel straight out of the matrix W and the last step writes it back.
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256)
panel_step_kernel(const float *__restrict__ in, long long ld_in, float *__restrict__ out, long long ld_out, int n,
                  int kb, int t, int k0, const u64 *__restrict__ part_in, int nparts, u64 *__restrict__ part_out,
                  float *__restrict__ CmT, long long ldc, int *__restrict__ piv, float *__restrict__ pv,
                  int *__restrict__ info, PanelState *__restrict__ ps) {
    __shared__ u64 sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = k0 + t;

    // ---- pivot of column t: every CTA reduces the partials redundantly
    u64 k = 0;
    for (int g = threadIdx.x; g < nparts; g += 256) {
        const u64 o = part_in[g];
        k = o > k ? o : k;
    }
    k = warp_max_u64(k);
    if (lane == 0) sm[warp] = k;
    __syncthreads();
    u64 best = sm[0];
#pragma unroll
    for (int w = 1; w < 8; w++) best = sm[w] > best ? sm[w] : best;
    __syncthreads();
    const int p = gj_key_row(best);
    const float v = gj_key_value(best);
    const float inv = 1.0f / v;

    const int tl = t >> 2, tc = t & 3;               // lane / component holding column t
    const int nl = (t + 1) >> 2, nc = (t + 1) & 3;   // ... column t+1
    const bool has_next = (t + 1) < kb;

    // ---- the two rows every CTA needs: pivot row (-> u) and the old row r (lands in row p)
    const float4 prow = *reinterpret_cast<const float4 *>(in + (long long)p * ld_in + lane * 4);
    const float4 rrow = *reinterpret_cast<const float4 *>(in + (long long)r * ld_in + lane * 4);
    float4 u;
    u.x = prow.x / v; u.y = prow.y / v; u.z = prow.z / v; u.w = prow.w / v;
    if (lane == tl) f4set(u, tc, inv);

    // ---- bookkeeping by CTA 0
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            piv[r] = p;
            pv[t] = v;
            if (gj_bad_pivot(v) && *info == 0) *info = r + 1;
        }
        // swaps also hit the multipliers already recorded for this panel (A.4 step 1)
        if (p != r && (int)threadIdx.x < t) {
            float *q = CmT + (long long)threadIdx.x * ldc;
            const float a = q[r], b = q[p];
            q[r] = b; q[p] = a;
        }
        // net row permutation of the panel, maintained incrementally for the row-block kernel
        if (warp == 7) {
            int m = (t == 0) ? kb : ps->m;
            if (t == 0)
                for (int i = lane; i < 2 * MATINV_NB; i += 32) { ps->pos[i] = k0 + i; ps->content[i] = i; }
            __syncwarp();
            if (p != r) {
                int b;
                if (p < k0 + kb) b = p - k0;
                else {
                    b = -1;
                    for (int base = kb; base < m; base += 32) {      // warp-parallel search of the outside slots
                        const int i = base + lane;
                        const bool hit = (i < m) && (ps->pos[i] == p);
                        const unsigned bal = __ballot_sync(0xffffffffu, hit);
                        if (bal) { b = base + (__ffs(bal) - 1); break; }
                    }
                    if (b < 0) {
                        b = m;
                        if (lane == 0) { ps->pos[m] = p; ps->content[m] = m; }
                        m++;
                        __syncwarp();
                    }
                }
                if (lane == 0) {
                    const int ca = ps->content[t], cb = ps->content[b];
                    ps->content[t] = cb; ps->content[b] = ca;
                }
            }
            if (lane == 0) ps->m = m;
        }
    }

    // ---- rank-1 update of my rows + partial pivot search of column t+1
    const int row_base = blockIdx.x * MATINV_RB + warp * 8;
    float4 src[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int i = row_base + q;
        if (i < n && i != r && i != p) src[q] = *reinterpret_cast<const float4 *>(in + (long long)i * ld_in + lane * 4);
        else src[q] = rrow;  // i == p receives the old row r; i == r / i >= n are overridden below
    }
    u64 mybest = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int i = row_base + q;
        if (i >= n) continue;                         // warp-uniform
        float4 o;
        float c = 0.0f;
        if (i == r) {
            o = u;
        } else {
            c = __shfl_sync(0xffffffffu, f4get(src[q], tc), tl);
            o.x = gj_elim(src[q].x, c, u.x);
            o.y = gj_elim(src[q].y, c, u.y);
            o.z = gj_elim(src[q].z, c, u.z);
            o.w = gj_elim(src[q].w, c, u.w);
            if (lane == tl) f4set(o, tc, fmaf(-c, inv, 0.0f));
        }
        *reinterpret_cast<float4 *>(out + (long long)i * ld_out + lane * 4) = o;
        if (lane == 0) CmT[(long long)t * ldc + i] = c;
        if (has_next && i > r && lane == nl) {
            const u64 kk = gj_key(f4get(o, nc), i, i == r + 1);
            mybest = kk > mybest ? kk : mybest;
        }
    }
    if (has_next) {
        mybest = __shfl_sync(0xffffffffu, mybest, nl);
        if (lane == 0) sm[warp] = mybest;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 b = sm[0];
#pragma unroll
            for (int w = 1; w < 8; w++) b = sm[w] > b ? sm[w] : b;
            part_out[blockIdx.x] = b;
        }
    }
}

void launch_panel_step(const float *in, long long ld_in, float *out, long long ld_out, int n, int kb, int t, int k0,
                       const u64 *part_in, int nparts, u64 *part_out, float *CmT, long long ldc, int *piv, float *pv,
                       int *info, PanelState *ps, cudaStream_t st) {
    panel_step_kernel<<<nparts, 256, 0, st>>>(in, ld_in, out, ld_out, n, kb, t, k0, part_in, nparts, part_out, CmT,
                                              ldc, piv, pv, info, ps);
}
