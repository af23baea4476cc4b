// Batched small-n Gauss-Jordan, v4: one warp per matrix, matrix resident in registers as packed pairs, BLOCKED in
// 8-column panels with LOOK-AHEAD (the A.4 scheme of the large-n path, SURVEY.md Appendix A.4, applied inside a warp).
//
// Replaces `for b: matrix_inv_32(A[b], n)` (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:317-362, five
// launches per column per matrix) for n in {32, 64}.  Same arithmetic as v3 (gj_batched_pk.cu) and the oracle -- every
// element sees the k-sequential chain a <- fma(-c, u, a) seeded from the element, pivot rows are scaled by IEEE
// division, the pivot column receives fma(-c, 1/v, +0) -- so results stay bit-identical.
//
// Why: ncu on v3 (profiles/r01_batched64_v3.*) shows one dependency chain per pivot step (search -> publish -> divide ->
// update -> reload) whose fixed ALU / shared-memory latencies nothing hides: two warps per scheduler, issue slots 42 %
// busy, top stall "wait".  Blocking alone does not help (measured: same rate); what helps is giving every warp
// INDEPENDENT work to issue in the shadow of its own chain:
//   (A) the 8-column panel g (window pairs 0..3) is factored step by step -- search, publish, divide, rank-1 update of
//       the panel columns only -- and between the dependent pieces of step t the warp issues pass t of the rank-8 update
//       that panel g-1 still owes the other columns: pure FFMA2 on registers + broadcast loads of U[t][.];
//   (B) the 8 pivot rows of panel g are brought up to date on the other columns by a COLUMN-parallel recurrence (lane =
//       column pair, the rows travel through shared memory), which also yields the snapshots U[t][.];
//   (C) the columns of panel g+1 (window pairs 4..7) receive panel g's rank-8 update at once, the rest is deferred to (A)
//       of the next panel;
//   (D) the owners of the 8 pivot rows reload them from shared memory (their registers took the update as garbage).
// Lane l owns physical rows l and l+32 (n = 64); row interchanges are implicit (logical positions); columns live in a
// window that rotates by 8 per panel so that every register index is static.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ u64 bk_pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void bk_unpack(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 bk_fma(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// 64-bit shared-memory accesses that the compiler must not merge into 128-bit ones (a merged access needs an aligned
// register quad and ptxas then gathers the pairs with moves)
__device__ __forceinline__ void sts64(unsigned addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 lds64(unsigned addr) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory"); return v; }

template <int N>
struct BlkLayout {
    static constexpr int RS = N / 32;            // rows per lane
    static constexpr int NP = N / 2;             // packed pairs per row
    static constexpr int PP = 4;                 // pairs of the 8-column panel
    static constexpr int OP = NP - PP;           // pairs of the other columns
    static constexpr int RP = NP - 2 * PP;       // pairs whose update is deferred to the next panel's factorisation
    static constexpr int PRB = OP * 8;           // bytes per published pivot row / per snapshot row U[t][.]
    static constexpr int PR_B = 8 * PRB;         // the 8 pivot rows of the panel (other columns)
    static constexpr int U_B = 8 * PRB;          // snapshots U[t][.]
    static constexpr int CMT_B = 8 * 8 * 4;      // multipliers of the 8 pivot rows, transposed: cmT[t][s]
    static constexpr int CMS_B = RS * 8 * 32 * 4;  // every row's multipliers of one panel: cmS[s][t][lane]; two buffers
    static constexpr int WORK_B = PR_B + U_B + CMT_B + 2 * CMS_B;
    static constexpr int OUT_B = 32 * (N + 1) * 4;   // output staging, 32 rows at a time
    static constexpr int MAIN_B = ((WORK_B > OUT_B ? WORK_B : OUT_B) + 15) / 16 * 16;
    static constexpr int MISC_B = 4 * 32 + N * 4 + 32 * 4;  // raw[8], u[8], v[8], prow[8], qinv[N], rowmap[32]
    static constexpr int PER_WARP_B = ((MAIN_B + MISC_B + 127) / 128) * 128;
    static_assert(RP >= 2 && RP % 4 == 0, "window layout");
};

template <int N, int WPC, int CPS>
__global__ void __launch_bounds__(32 * WPC, CPS)
batched_blk_kernel(const float *__restrict__ A, long long batch, float *__restrict__ X, int *__restrict__ info) {
    typedef BlkLayout<N> L;
    constexpr int RS = L::RS, NP = L::NP, PP = L::PP, OP = L::OP, RP = L::RP, PRB = L::PRB;
    constexpr int LD = N + 1;
    extern __shared__ __align__(16) unsigned char smem_blk[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_blk + warp * L::PER_WARP_B;
    unsigned char *PR = base;
    unsigned char *Ub = base + L::PR_B;
    float *cmT = reinterpret_cast<float *>(base + L::PR_B + L::U_B);
    float *cmS = reinterpret_cast<float *>(base + L::PR_B + L::U_B + L::CMT_B);
    float *raw = reinterpret_cast<float *>(base + L::MAIN_B);
    float *uu = raw + 8, *vs = raw + 16;
    int *prow = reinterpret_cast<int *>(raw + 24);
    int *qinv = reinterpret_cast<int *>(raw + 32);
    int *rowmap = qinv + N;
    float *ostage = reinterpret_cast<float *>(base);  // 32 x (n+1) floats, output staging only
    const unsigned raw_s = (unsigned)__cvta_generic_to_shared(raw);
    const unsigned PR_s = (unsigned)__cvta_generic_to_shared(PR);
    const int qB = lane % OP;     // column pair of this lane in (B); lanes >= OP repeat a pair (same values, same addresses)

    for (long long b = (long long)blockIdx.x * WPC + warp; b < batch; b += (long long)gridDim.x * WPC) {
        const float *Ab = A + b * (long long)(N * N);
        u64 a2[RS][NP];
        int lpos[RS];
#pragma unroll
        for (int s = 0; s < RS; s++) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(Ab + (s * 32 + lane) * N);
#pragma unroll
            for (int f = 0; f < N / 4; f++) {
                const ulonglong2 v4 = src[f];
                a2[s][2 * f] = v4.x;
                a2[s][2 * f + 1] = v4.y;
            }
            lpos[s] = s * 32 + lane;
            qinv[s * 32 + lane] = s * 32 + lane;
        }
        // the first panel has no deferred update: zero multipliers and snapshots make its passes exact no-ops
        // (fma(-0, 0, a) = a for every a, signed zeros included)
#pragma unroll
        for (int i = lane; i < L::CMS_B / 4; i += 32) cmS[L::CMS_B / 4 + i] = 0.0f;
#pragma unroll
        for (int i = lane; i < L::U_B / 4; i += 32) reinterpret_cast<float *>(Ub)[i] = 0.0f;
        int sinfo = 0;
        int psidx[RS];     // pivot step of this row inside the PREVIOUS panel (its deferred reload), else out of range
#pragma unroll
        for (int s = 0; s < RS; s++) psidx[s] = -1;
        __syncwarp();

        // pass t of the deferred rank-8 update of the previous panel on window pairs [j0, j1) (both even)
        auto deferred_pass = [&](const float *cmPrev, int t, int j0, int j1) {
            u64 ncm[RS];
#pragma unroll
            for (int s = 0; s < RS; s++) {
                const float c = cmPrev[(s * 8 + t) * 32 + lane];
                ncm[s] = bk_pack(-c, -c);
            }
            const ulonglong2 *ut = reinterpret_cast<const ulonglong2 *>(Ub + t * PRB);
#pragma unroll
            for (int j = j0; j < j1; j += 2) {
                const ulonglong2 u4 = ut[j / 2];      // U index of window pair j is j (the window moved by PP since U was written)
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    a2[s][j] = bk_fma(ncm[s], u4.x, a2[s][j]);
                    a2[s][j + 1] = bk_fma(ncm[s], u4.y, a2[s][j + 1]);
                }
            }
        };
        auto deferred_reload = [&]() {
#pragma unroll
            for (int s = 0; s < RS; s++) {
                if ((unsigned)psidx[s] < 8u) {
                    const unsigned src = PR_s + psidx[s] * PRB;
#pragma unroll
                    for (int j = PP; j < PP + RP; j++) a2[s][j] = lds64(src + 8 * j);
                }
            }
        };
        constexpr int JH = PP + RP / 2;   // the deferred pairs [PP, PP+RP) are issued in two pieces per step

#pragma unroll 1
        for (int g = 0; g < N / 8; g++) {
            float *cmCur = cmS + (g & 1) * (L::CMS_B / 4);
            const float *cmPrev = cmS + ((g & 1) ^ 1) * (L::CMS_B / 4);
            // ================= (A) factor the 8-column panel (pairs 0..3), deferred passes in the shadows =================
#pragma unroll
            for (int tc = 0; tc < 8; tc++) {
                const int r = 8 * g + tc;
                // ---- arg max over the unspent rows (logical position >= r), lowest position on ties
                float cm[RS];
                unsigned mag = 0;
                unsigned cand = 0x7FFFFFFFu;
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    float lo, hi;
                    bk_unpack(a2[s][tc >> 1], lo, hi);
                    cm[s] = (tc & 1) ? hi : lo;
                    const unsigned m = __float_as_uint(cm[s]) & 0x7FFFFFFFu;
                    const bool live = lpos[s] >= r;
                    const bool isnum = m <= 0x7F800000u;
                    // NaN: a candidate never wins (key 0), the incumbent (position r) is never displaced (key all ones)
                    const unsigned mq = live ? (isnum ? m : (lpos[s] == r ? 0xFFFFFFFFu : 0u)) : 0u;
                    const unsigned lq = live ? (unsigned)lpos[s] : 0x7FFFFFFFu;
                    if (s == 0 || mq > mag || (mq == mag && lq < cand)) { mag = mq; cand = lq; }
                }
                const unsigned gm = __reduce_max_sync(0xffffffffu, mag);
                const int p = (int)__reduce_min_sync(0xffffffffu, mag == gm ? cand : 0x7FFFFFFFu);
                bool own[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) own[s] = lpos[s] == p;
                // ---- the owner publishes the raw panel row (8 floats) and where the row lives
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    if (own[s]) {
#pragma unroll
                        for (int j = 0; j < PP; j++) sts64(raw_s + 8 * j, a2[s][j]);
                        prow[tc] = s * 32 + lane;
                    }
                }
                __syncwarp();
                deferred_pass(cmPrev, tc, PP, JH);
                // ---- true division, one element per lane (lanes >= 8 repeat); the pivot position receives 1/v
                const float v = raw[tc];
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
                const float num = ((lane & 7) == tc) ? 1.0f : raw[lane & 7];
                uu[lane & 7] = __fdiv_rn(num, v);
                vs[tc] = v;
                __syncwarp();
                deferred_pass(cmPrev, tc, JH, PP + RP);
                u64 u2[PP];
                {
                    const ulonglong2 ua = reinterpret_cast<const ulonglong2 *>(uu)[0];
                    const ulonglong2 ub = reinterpret_cast<const ulonglong2 *>(uu)[1];
                    u2[0] = ua.x; u2[1] = ua.y; u2[2] = ub.x; u2[3] = ub.y;
                }
                // ---- rank-1 update of the panel columns; the pivot column starts from +0 so that it receives -c/v
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    float lo, hi;
                    bk_unpack(a2[s][tc >> 1], lo, hi);
                    a2[s][tc >> 1] = (tc & 1) ? bk_pack(lo, 0.0f) : bk_pack(0.0f, hi);
                    cmCur[(s * 8 + tc) * 32 + lane] = cm[s];
                    const u64 ncm = bk_pack(-cm[s], -cm[s]);
#pragma unroll
                    for (int j = 0; j < PP; j++) {
                        const u64 upd = bk_fma(ncm, u2[j], a2[s][j]);
                        a2[s][j] = own[s] ? u2[j] : upd;   // the pivot row becomes u
                    }
                }
                // ---- bookkeeping: logical positions; lane 0 keeps the column permutation
#pragma unroll
                for (int s = 0; s < RS; s++) lpos[s] = own[s] ? r : (lpos[s] == r ? p : lpos[s]);
                if (lane == 0) {
                    const int q1 = qinv[r], q2 = qinv[p];
                    qinv[r] = q2;
                    qinv[p] = q1;
                }
            }
            // the previous panel's pivot rows: their deferred columns come from the stage (written by (B) of that panel)
            deferred_reload();
            __syncwarp();

            // ================= (B) the 8 pivot rows on the other columns, column-parallel =================
            // The owner of a pivot row publishes the row (other columns) and its 8 multipliers, indexed by pivot step.
            int sidx[RS];
#pragma unroll
            for (int s = 0; s < RS; s++) {
                sidx[s] = lpos[s] - 8 * g;                 // pivot step of this row inside the panel, if any
                if ((unsigned)sidx[s] < 8u) {
                    float mine[8];     // all loads first: the compiler cannot prove cmS and cmT distinct and would chain load -> store
#pragma unroll
                    for (int t = 0; t < 8; t++) mine[t] = cmCur[(s * 8 + t) * 32 + lane];
#pragma unroll
                    for (int t = 0; t < 8; t++) cmT[8 * t + sidx[s]] = mine[t];
                    const unsigned dst = PR_s + sidx[s] * PRB;
#pragma unroll
                    for (int j = 0; j < OP; j++) sts64(dst + 8 * j, a2[s][PP + j]);
                }
            }
            __syncwarp();
            {
                u64 x[8];
#pragma unroll
                for (int s = 0; s < 8; s++) x[s] = *reinterpret_cast<const u64 *>(PR + s * PRB + qB * 8);
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const float vt = vs[t];
                    float lo, hi;
                    bk_unpack(x[t], lo, hi);
                    const u64 ut = bk_pack(__fdiv_rn(lo, vt), __fdiv_rn(hi, vt));
                    x[t] = ut;
                    *reinterpret_cast<u64 *>(Ub + t * PRB + qB * 8) = ut;
                    const float4 c0 = reinterpret_cast<const float4 *>(cmT + 8 * t)[0];
                    const float4 c1 = reinterpret_cast<const float4 *>(cmT + 8 * t)[1];
                    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int s = 0; s < 8; s++)
                        if (s != t) x[s] = bk_fma(bk_pack(-cc[s], -cc[s]), ut, x[s]);
                }
#pragma unroll
                for (int s = 0; s < 8; s++) *reinterpret_cast<u64 *>(PR + s * PRB + qB * 8) = x[s];
            }
            __syncwarp();

            // ================= (C) the next panel's columns (pairs 4..7) take the rank-8 update now =================
#pragma unroll
            for (int t = 0; t < 8; t++) {
                u64 ncm[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    const float c = cmCur[(s * 8 + t) * 32 + lane];
                    ncm[s] = bk_pack(-c, -c);
                }
                const ulonglong2 *ut = reinterpret_cast<const ulonglong2 *>(Ub + t * PRB);
#pragma unroll
                for (int k = 0; k < PP / 2; k++) {
                    const ulonglong2 u4 = ut[k];
#pragma unroll
                    for (int s = 0; s < RS; s++) {
                        a2[s][PP + 2 * k] = bk_fma(ncm[s], u4.x, a2[s][PP + 2 * k]);
                        a2[s][PP + 2 * k + 1] = bk_fma(ncm[s], u4.y, a2[s][PP + 2 * k + 1]);
                    }
                }
            }
            // ================= (D) the owners reload the finished pivot rows on those columns =================
#pragma unroll
            for (int s = 0; s < RS; s++) {
                if ((unsigned)sidx[s] < 8u) {
                    const unsigned src = PR_s + sidx[s] * PRB;
#pragma unroll
                    for (int j = 0; j < PP; j++) a2[s][PP + j] = lds64(src + 8 * j);
                }
                psidx[s] = sidx[s];
            }
            // ---- rotate the window left by 8 columns: the next panel becomes pairs 0..3, this one goes to the end
#pragma unroll
            for (int s = 0; s < RS; s++) {
                u64 tmp[PP];
#pragma unroll
                for (int j = 0; j < PP; j++) tmp[j] = a2[s][j];
#pragma unroll
                for (int j = 0; j < OP; j++) a2[s][j] = a2[s][j + PP];
#pragma unroll
                for (int j = 0; j < PP; j++) a2[s][OP + j] = tmp[j];
            }
            __syncwarp();
        }
        // ---- the last panel's deferred update (window pairs 4..NP-5 after the final rotation)
        {
            const float *cmPrev = cmS + (((N / 8) & 1) ^ 1) * (L::CMS_B / 4);
#pragma unroll
            for (int t = 0; t < 8; t++) deferred_pass(cmPrev, t, PP, PP + RP);
            deferred_reload();
        }

        // ---- result: X[lpos][qinv[col]] = a[.][col]; after N/8 rotations the window is back at column 0.  Staged through
        //      shared memory 32 rows at a time for coalesced stores.
        __syncwarp();
        float *Xb = X + b * (long long)(N * N);
        bool bad = false;
#pragma unroll
        for (int s = 0; s < RS; s++) {
            float *row = ostage + lane * LD;
#pragma unroll
            for (int P = 0; P < NP; P++) {
                float lo, hi;
                bk_unpack(a2[s][P], lo, hi);
                row[qinv[2 * P]] = lo;
                row[qinv[2 * P + 1]] = hi;
            }
            rowmap[lane] = lpos[s];
            __syncwarp();
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const int grow = rowmap[i];
#pragma unroll
                for (int k = 0; k < RS; k++) {
                    const float xv = ostage[i * LD + lane + 32 * k];
                    bad |= !isfinite(xv);
                    Xb[grow * N + lane + 32 * k] = xv;
                }
            }
            __syncwarp();
        }
        const int anybad = __any_sync(0xffffffffu, bad);
        if (lane == 0 && info) info[b] = sinfo ? sinfo : (anybad ? -1 : 0);
        __syncwarp();
    }
}

template <int N, int WPC, int CPS>
cudaError_t launch_blk(const float *A, long long batch, float *X, int *info, cudaStream_t st) {
    const size_t smem = (size_t)WPC * BlkLayout<N>::PER_WARP_B;
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(batched_blk_kernel<N, WPC, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    long long grid = (batch + WPC - 1) / WPC;
    const long long cap = 148ll * CPS * 8;
    if (grid > cap) grid = cap;
    batched_blk_kernel<N, WPC, CPS><<<(unsigned)grid, 32 * WPC, smem, st>>>(A, batch, X, info);
    return cudaGetLastError();
}

}  // namespace

// n in {32, 64}.  n = 64: 4 warps per CTA; MATINV_BLK_CPS = 2 (255 registers, two matrices per scheduler) or 3 (168
// registers, three matrices per scheduler).
cudaError_t launch_batched_blk(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st) {
    static int cps = -1;
    if (cps < 0) {
        const char *e = getenv("MATINV_BLK_CPS");
        cps = e ? atoi(e) : 2;
    }
    if (n == 64) return cps == 3 ? launch_blk<64, 4, 3>(A, batch, X, info, st) : launch_blk<64, 4, 2>(A, batch, X, info, st);
    if (n == 32) return launch_blk<32, 4, 4>(A, batch, X, info, st);
    return cudaErrorInvalidValue;
}
