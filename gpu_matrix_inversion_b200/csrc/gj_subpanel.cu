// Panel factorisation, v1: the latency-critical part of the blocked Gauss-Jordan.
//
// A 128-wide panel is factored as W-wide sub-panels (W = 16, or 8 for very tall panels):
//
//   subpanel_kernel      ONE thread-block cluster (up to 16 CTAs x 512 threads) holds the whole
//                        n x W sub-panel in REGISTERS and runs its W pivot steps without leaving the
//                        chip: per step a warp-shuffle arg max, one DSMEM all-to-all of
//                        (key, candidate row) between the CTAs, one cluster barrier, and the rank-1
//                        update from registers.  Fuses, per column, the reference's maxPivotKernel +
//                        finalMaxPivotKernel + pivotElementsKernel + fixRowKernel + fixColumnKernel
//                        (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:13-173, launch
//                        loop :317-362) restricted to the sub-panel columns.
//   panel_update_kernel  applies the sub-panel's W steps to the other columns of the panel (whole
//                        GPU): row interchanges, recurrence on the W pivot rows, rank-W update.
//
// Row interchanges inside the cluster kernel are IMPLICIT: a row never leaves its registers, each
// row carries its logical position `lpos` and the swap r <-> p just exchanges two logical positions.
// Candidates are rows with lpos >= r and ties break on the lowest lpos, so the pivot sequence and
// every FMA chain are exactly those of the physically swapping algorithm (oracle A.3/A.4).
//
// The panel ping-pongs between two buffers per sub-panel (like the reference's two [A|I] buffers,
// :352-359), so the update kernel never reads a row another CTA has already rewritten.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define SP_THREADS 512
#define SP_MAXCTA 16

// Optional in-kernel timeline (SM clock) of the LAST launch, for tuning: matinv_debug_trace().
__device__ int g_trace_on = 0;
__device__ long long g_trace[128];
#define TRACE(cond, slot) do { if (trace_on && (cond) && threadIdx.x == 0) g_trace[slot] = clock64(); } while (0)

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa_shared(const void *p, unsigned cta) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(unsigned addr, const float4 &v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void st_async_f4(unsigned addr, const float4 &v, unsigned mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(a), "r"(parity)
        : "memory");
}

struct __align__(16) SubMail {  // 80 bytes = 5 float4
    u64 key;
    u64 pad;
    float row[16];
};

struct __align__(16) SubSmem {
    SubMail mail[2][SP_MAXCTA];       // [step parity][source CTA]  -- written remotely (st.async)
    SubMail wcand[SP_THREADS / 32];   // staging of the CTA winner's candidate (only the winner's warp uses its slot)
    float uw[SP_THREADS / 32][16];    // per-warp copy of the normalised pivot row (rotated)
    u64 mbar[2];                      // one transaction barrier per step parity
    u64 cta_key[3];                   // CTA-level arg max by shared-memory atomicMax, reset two steps ahead
    u64 pad;
    // followed by: float hist[W][R * SP_THREADS]
};

// ------------------------------------------------------------------------------------------ sub-panel
template <int W, int R, int TH>
__global__ void __launch_bounds__(TH, 1)
subpanel_kernel(const float *__restrict__ in, long long ld_in, float *__restrict__ out, long long ld_out, int n, int k0,
                int s0, int sw, float *__restrict__ CmT, long long ldc, int *__restrict__ piv, float *__restrict__ pv,
                int *__restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SubSmem &s = *reinterpret_cast<SubSmem *>(smem_raw);
    float *hist = reinterpret_cast<float *>(smem_raw + sizeof(SubSmem));  // [W][R*TH]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster_ctarank(), nct = cluster_nctarank();
    const int trace_on = g_trace_on;
    TRACE(rank == 0, 0);

    float x[R][W];
    int lpos[R];
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int i = (q * (int)nct + (int)rank) * TH + tid;
        lpos[q] = (i < n) ? i : -1;
        if (i < n) {
            const float4 *src = reinterpret_cast<const float4 *>(in + (long long)i * ld_in + s0);
#pragma unroll
            for (int f = 0; f < W / 4; f++) {
                const float4 v4 = src[f];
                x[q][4 * f + 0] = v4.x; x[q][4 * f + 1] = v4.y; x[q][4 * f + 2] = v4.z; x[q][4 * f + 3] = v4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) x[q][j] = 0.0f;
        }
    }
    TRACE(rank == 0, 1);
    if (tid == 0) {
        mbar_init(&s.mbar[0], 1);
        mbar_init(&s.mbar[1], 1);
        s.cta_key[0] = 0; s.cta_key[1] = 0; s.cta_key[2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // every CTA of the cluster must be resident (and its barriers initialised) before anyone writes into its
    // shared memory
    cluster_arrive();
    cluster_wait();
    TRACE(rank == 0, 2);

    // The step loop is fully UNROLLED: every register index (current column t) is static.  (A rolled loop with
    // a rotating register window was tried: ptxas turns the rotation into ~300 register moves per step.)
#pragma unroll
    for (int t = 0; t < W; t++) {
        if (t < sw) {
            const int r = k0 + s0 + t;
            TRACE(rank == 0 && t == 8, 64);
            // ---- (1) local candidates -> warp arg max (two redux.sync) -> CTA arg max (smem atomicMax)
            const int par = t & 1;
            if (tid == 0) mbar_expect_tx(&s.mbar[par], nct * (16 + 4 * W));
            unsigned mag = 0;
            int cand = 0x7FFFFFFF, bq = -1;
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (lpos[q] >= r) {
                    const unsigned mq = gj_mag(x[q][t], lpos[q] == r);
                    if (bq < 0 || mq > mag || (mq == mag && lpos[q] < cand)) { mag = mq; cand = lpos[q]; bq = q; }
                }
            }
            const bool has = bq >= 0;
            const unsigned gm = __reduce_max_sync(0xffffffffu, has ? mag : 0u);
            const unsigned pw = __reduce_min_sync(0xffffffffu, (has && mag == gm) ? (unsigned)cand : 0x7FFFFFFFu);
            u64 mykey = 0;
            if (has && mag == gm && (unsigned)cand == pw) {  // exactly one lane per warp: positions are distinct
                float xv = 0.0f;
#pragma unroll
                for (int q = 0; q < R; q++)
                    if (q == bq) xv = x[q][t];
                mykey = gj_key_from(mag, cand, xv);
                atomicMax(&s.cta_key[t % 3], mykey);
            }
            TRACE(rank == 0 && t == 8, 65);
            __syncthreads();
            TRACE(rank == 0 && t == 8, 66);
            if (tid == 0) s.cta_key[(t + 2) % 3] = 0;
            // ---- (2) the winner's warp pushes (key, row) into every CTA's mailbox: DSMEM st.async, completion
            //          counted by the destination's transaction barrier -- no cluster-wide barrier per step
            const u64 ck = s.cta_key[t % 3];
            const bool iwin = (mykey != 0) && (mykey == ck);
            const bool nocand = (ck == 0);  // no candidate row in this CTA: warp 0 sends the empty key
            if (__any_sync(0xffffffffu, iwin) || (nocand && warp == 0)) {
                SubMail &c = s.wcand[warp];
                if (iwin) {
                    c.key = mykey;
#pragma unroll
                    for (int q = 0; q < R; q++)
                        if (q == bq) {
#pragma unroll
                            for (int f = 0; f < W / 4; f++)
                                *reinterpret_cast<float4 *>(&c.row[4 * f]) =
                                    make_float4(x[q][4 * f], x[q][4 * f + 1], x[q][4 * f + 2], x[q][4 * f + 3]);
                        }
                } else if (nocand && lane == 0) {
                    c.key = 0;
                }
                __syncwarp();
                const float4 *src = reinterpret_cast<const float4 *>(&c);
                const unsigned dst_cta = lane & 15;
                if (dst_cta < nct) {
                    const unsigned base = mapa_shared(&s.mail[par][rank], dst_cta);
                    const unsigned rbar = mapa_shared(&s.mbar[par], dst_cta);
                    for (int f = lane >> 4; f < 1 + W / 4; f += 2) st_async_f4(base + 16 * f, src[f], rbar);
                }
            }
            TRACE(rank == 0 && t == 8, 67);
            mbar_wait(&s.mbar[par], (t >> 1) & 1);
            TRACE(rank == 0 && t == 8, 68);
            // ---- (3) every warp: cluster arg max -> pivot row p, value v, normalised pivot row u (rotated)
            int p;
            {
                const bool inr = lane < (int)nct;
                const u64 k = inr ? s.mail[par][lane].key : 0;
                const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
                const unsigned ghi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned glo = __reduce_max_sync(0xffffffffu, (inr && hi == ghi) ? lo : 0u);
                const unsigned hit = __ballot_sync(0xffffffffu, inr && hi == ghi && lo == glo);
                const int csrc = __ffs(hit) - 1;
                const u64 kg = ((u64)ghi << 32) | glo;
                p = gj_key_row(kg);
                const float v = gj_key_value(kg);
                if (lane < W) {
                    const float rowv = s.mail[par][csrc].row[lane];
                    s.uw[warp][lane] = (lane == t) ? 1.0f / v : rowv / v;
                }
                if (rank == 0 && tid == 0) {
                    piv[r] = p;
                    pv[s0 + t] = v;
                    if (gj_bad_pivot(v) && *info == 0) *info = r + 1;
                }
            }
            __syncwarp();
            TRACE(rank == 0 && t == 8, 69);
            float u[W];
#pragma unroll
            for (int f = 0; f < W / 4; f++) {
                const float4 v4 = *reinterpret_cast<const float4 *>(&s.uw[warp][4 * f]);
                u[4 * f] = v4.x; u[4 * f + 1] = v4.y; u[4 * f + 2] = v4.z; u[4 * f + 3] = v4.w;
            }
            __syncwarp();
            TRACE(rank == 0 && t == 8, 70);
            // ---- (4) implicit swap + rank-1 update from registers; multipliers recorded in smem
#pragma unroll
            for (int q = 0; q < R; q++) {
                if (lpos[q] < 0) continue;
                float c = 0.0f;
                if (lpos[q] == p) {
                    lpos[q] = r;
#pragma unroll
                    for (int j = 0; j < W; j++) x[q][j] = u[j];
                } else {
                    if (lpos[q] == r) lpos[q] = p;
                    c = x[q][t];
#pragma unroll
                    for (int j = 0; j < W; j++)
                        if (j != t) x[q][j] = gj_elim(x[q][j], c, u[j]);
                    x[q][t] = fmaf(-c, u[t], 0.0f);
                }
                hist[t * (R * TH) + q * TH + tid] = c;
            }
            TRACE(rank == 0, 3 + t);
        }
    }

    // ---- write back at the logical positions
#pragma unroll
    for (int q = 0; q < R; q++) {
        const int i = lpos[q];
        if (i < 0) continue;
        float4 *dst = reinterpret_cast<float4 *>(out + (long long)i * ld_out + s0);
#pragma unroll
        for (int f = 0; f < W / 4; f++) dst[f] = make_float4(x[q][4 * f], x[q][4 * f + 1], x[q][4 * f + 2], x[q][4 * f + 3]);
#pragma unroll
        for (int t = 0; t < W; t++)
            if (t < sw) CmT[(long long)(s0 + t) * ldc + i] = hist[t * (R * TH) + q * TH + tid];
    }
    TRACE(rank == 0, 20);
    // nobody may exit while a peer could still write into its mailbox
    cluster_arrive();
    cluster_wait();
    TRACE(rank == 0, 21);
}

// ------------------------------------------------------------------------------------------ update of the rest of the panel
template <int ROWS>
struct __align__(16) UpdSmem {
    float old_[32][MATINV_NB];   // original contents of every row touched by the sub-panel's swaps
    float us[16][MATINV_NB];     // U snapshot of the sub-panel steps
    float xf[16][MATINV_NB];     // final contents of the sub-panel's pivot rows
    float cs[16][ROWS];          // multipliers of my rows
    float cp[16][16];            // multipliers of the pivot rows
    float pv[16];
    int pos[32], content[32];
    int m;
    int rowmap[ROWS];            // my row -> slot in pos[] or -1
};

// Net permutation of the sub-panel's sw (<= 16) row interchanges, one warp, no serial search: every lane tracks ONE
// row through the sw transpositions.  Slots 0..15 are the pivot rows r0+t, slots 16..31 the distinct outside rows
// that were picked as pivots (slot 16+t = p_t at its first occurrence).  Output, 32 entries each:
//   pos[idx]     row of slot idx, -1 if the slot is unused
//   content[idx] slot whose ORIGINAL data ends up in slot idx
__device__ __forceinline__ void build_subperm(const int *__restrict__ piv, int r0, int sw, int *pos, int *content, int *mout) {
    const int lane = threadIdx.x & 31;
    const int pl = (lane < sw) ? piv[r0 + lane] : -1;     // p_t for t = lane (one coalesced read)
    // all 32 lanes execute the collectives; lanes < 16 carry unique dummy keys through match_any
    const int cand = __shfl_sync(0xffffffffu, pl, (lane - 16) & 31);
    const bool outside = lane >= 16 && cand >= r0 + sw;
    const unsigned grp = __match_any_sync(0xffffffffu, outside ? cand : -1 - lane);
    int row;
    if (lane < 16) row = (lane < sw) ? r0 + lane : -1;
    else row = (outside && lane == __ffs(grp) - 1) ? cand : -1;
    int where = row;                                       // where this slot's original data currently sits
    for (int t = 0; t < sw; t++) {
        const int pt = __shfl_sync(0xffffffffu, pl, t), rt = r0 + t;
        if (where >= 0 && pt != rt) {
            if (where == rt) where = pt;
            else if (where == pt) where = rt;
        }
    }
    pos[lane] = row;
    __syncwarp();
    if (row >= 0) {   // my data ends in the slot whose row is `where`
        int k = -1;
        if (where < r0 + sw) k = where - r0;
        else {
            for (int q = 16; q < 32; q++)
                if (pos[q] == where) k = q;
        }
        content[k] = lane;
    }
    __syncwarp();
    if (lane == 0) *mout = 32;
}

// ROWS rows per CTA.  Everything before the main loop (permutation, gather of the touched rows, 16-step recurrence on
// the pivot rows) is the same for every CTA and is a latency chain of ~12k cycles, the update itself is ~1.5k cycles
// per 64 rows: with 256 rows per CTA the kernel holds a quarter of the SM slots for about the same time, which is what
// matters when it runs beside the trailing update (look-ahead): its CTAs displace GEMM CTAs for as long as they live.
template <int ROWS>
__global__ void __launch_bounds__(256, 2)   // <= 128 registers: a CTA must fit beside one CTA of the trailing update
panel_update_kernel(const float *__restrict__ in, long long ld_in, float *__restrict__ out, long long ld_out, int n,
                    int k0, int s0, int sw, int wfull, float *__restrict__ CmT, long long ldc,
                    const int *__restrict__ piv, const float *__restrict__ pvg, PanelState *__restrict__ ps, int kb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    UpdSmem<ROWS> &s = *reinterpret_cast<UpdSmem<ROWS> *>(smem_raw);
    constexpr int RPW = ROWS / 8;   // rows per warp
    constexpr int GR = 4;           // rows per group (their 16-step chains are interleaved)
    constexpr int NG = RPW / GR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = k0 + s0;
    const int trace_on = g_trace_on;
    TRACE(blockIdx.x == 3, 32);
    TRACE(blockIdx.x == gridDim.x - 1, 48);

    if (blockIdx.x == gridDim.x - 1) {
        // ===== bookkeeping CTA: whole-panel permutation state + swaps of the earlier multiplier columns
        int *ppos = reinterpret_cast<int *>(smem_raw);               // [256]
        int *pcontent = ppos + 2 * MATINV_NB;                         // [256]
        int *spos = pcontent + 2 * MATINV_NB;                         // [32]
        int *scontent = spos + 32;                                    // [32]
        int *sm_m = scontent + 32;                                    // [2]
        float *oldh = reinterpret_cast<float *>(sm_m + 4);            // [32][s0]
        for (int i = tid; i < 2 * MATINV_NB; i += 256) {
            ppos[i] = (s0 == 0) ? k0 + i : ps->pos[i];
            pcontent[i] = (s0 == 0) ? i : ps->content[i];
        }
        __syncthreads();
        if (warp == 0) {
            int m = (s0 == 0) ? kb : ps->m;
            const int pl = (lane < sw) ? piv[r0 + lane] : 0;
            for (int t = 0; t < sw; t++) {
                const int r = r0 + t, p = __shfl_sync(0xffffffffu, pl, t);
                if (p == r) continue;
                int b = -1;
                if (p < k0 + kb) b = p - k0;
                else {
                    for (int base = kb; base < m; base += 32) {
                        const int i = base + lane;
                        const bool hit = (i < m) && (ppos[i] == p);
                        const unsigned bal = __ballot_sync(0xffffffffu, hit);
                        if (bal) { b = base + (__ffs(bal) - 1); break; }
                    }
                    if (b < 0) {
                        b = m;
                        if (lane == 0) { ppos[m] = p; pcontent[m] = m; }
                        m++;
                    }
                }
                __syncwarp();
                if (lane == 0) { const int ca = pcontent[s0 + t], cb = pcontent[b]; pcontent[s0 + t] = cb; pcontent[b] = ca; }
                __syncwarp();
            }
            if (lane == 0) ps->m = m;
        } else if (warp == 1) {
            build_subperm(piv, r0, sw, spos, scontent, sm_m);
        }
        __syncthreads();
        for (int i = tid; i < 2 * MATINV_NB; i += 256) { ps->pos[i] = ppos[i]; ps->content[i] = pcontent[i]; }
        const int m = sm_m[0];
        TRACE(true, 49);
        // multipliers recorded by the earlier sub-panels of this panel follow their rows
        for (int e = tid; e < m * s0; e += 256) {
            const int idx = e / s0, q = e - idx * s0;
            if (spos[idx] >= 0) oldh[idx * s0 + q] = CmT[(long long)q * ldc + spos[idx]];
        }
        __syncthreads();
        for (int e = tid; e < m * s0; e += 256) {
            const int idx = e / s0, q = e - idx * s0;
            const int c = scontent[idx];
            if (spos[idx] >= 0 && c != idx) CmT[(long long)q * ldc + spos[idx]] = oldh[c * s0 + q];
        }
        TRACE(true, 50);
        return;
    }

    // ===== regular CTA: rows [i0, i0 + ROWS), warp w owns rows i0 + w*RPW + [0, RPW) in groups of GR
    const int i0 = blockIdx.x * ROWS;
    // the first group of rows is fetched up front: the loads fly while the permutation / recurrence prologue runs
    float4 pre[GR];
#pragma unroll
    for (int q = 0; q < GR; q++) {
        const int i = i0 + warp * RPW + q;
        pre[q] = (i < n) ? *reinterpret_cast<const float4 *>(in + (long long)i * ld_in + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (warp == 0) build_subperm(piv, r0, sw, s.pos, s.content, &s.m);
    for (int e = tid; e < 16 * ROWS; e += 256) {
        const int t = e / ROWS, ii = e - t * ROWS;
        s.cs[t][ii] = (t < sw && i0 + ii < n) ? CmT[(long long)(s0 + t) * ldc + i0 + ii] : 0.0f;
    }
    for (int ii = tid; ii < ROWS; ii += 256) s.rowmap[ii] = -1;
    {
        const int t = tid >> 4, t2 = tid & 15;
        s.cp[t][t2] = (t < sw && t2 < sw) ? CmT[(long long)(s0 + t) * ldc + r0 + t2] : 0.0f;
        if (tid < 16) s.pv[tid] = (tid < sw) ? pvg[s0 + tid] : 1.0f;
    }
    __syncthreads();
    TRACE(blockIdx.x == 3, 33);
    const int m = s.m;
    if (tid < m && s.pos[tid] >= 0) {
        const int ii = s.pos[tid] - i0;
        if (ii >= 0 && ii < ROWS) s.rowmap[ii] = tid;
    }
    for (int e = tid; e < m * 32; e += 256) {
        const int idx = e >> 5, f = e & 31;
        if (s.pos[idx] >= 0)
            *reinterpret_cast<float4 *>(&s.old_[idx][4 * f]) =
                *reinterpret_cast<const float4 *>(in + (long long)s.pos[idx] * ld_in + 4 * f);
    }
    __syncthreads();
    TRACE(blockIdx.x == 3, 34);
    // recurrence on the sw pivot rows, one thread per panel column; fully unrolled with static indices so that the
    // reciprocal part of every division (it depends on pv only) is scheduled off the dependent chain
    if (tid < MATINV_NB) {
        float xx[16];
#pragma unroll
        for (int t = 0; t < 16; t++) xx[t] = (t < sw) ? s.old_[s.content[t]][tid] : 0.0f;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            if (t < sw) {
                const float u = xx[t] / s.pv[t];
                s.us[t][tid] = u;
                xx[t] = u;
#pragma unroll
                for (int t2 = 0; t2 < 16; t2++)
                    if (t2 != t) xx[t2] = gj_elim(xx[t2], s.cp[t][t2], u);   // cp is 0 beyond sw: exact no-op
            } else {
                s.us[t][tid] = 0.0f;
            }
        }
#pragma unroll
        for (int t = 0; t < 16; t++) s.xf[t][tid] = xx[t];
    }
    __syncthreads();

    TRACE(blockIdx.x == 3, 35);
    const bool own_cols = (lane >= (s0 >> 2)) && (lane < ((s0 + wfull) >> 2));  // the sub-panel's own columns
    float4 us[16];
#pragma unroll
    for (int t = 0; t < 16; t++) us[t] = *reinterpret_cast<const float4 *>(&s.us[t][4 * lane]);
#pragma unroll 2
    for (int g = 0; g < NG; g++) {
        const int ii0 = warp * RPW + g * GR;
        // GR rows at a time, the 16-step chains interleaved: rows touched by a swap start from their permuted contents,
        // the sub-panel's own pivot rows take the recurrence result afterwards (their chain is discarded)
        float4 acc[GR];
#pragma unroll
        for (int q = 0; q < GR; q++) {
            const int slot = s.rowmap[ii0 + q];
            acc[q] = pre[q];
            if (slot >= 0) acc[q] = *reinterpret_cast<const float4 *>(&s.old_[s.content[slot]][4 * lane]);
        }
        if (g + 1 < NG) {   // next group's rows fly during this group's FMAs
#pragma unroll
            for (int q = 0; q < GR; q++) {
                const int i = i0 + ii0 + GR + q;
                pre[q] = (i < n) ? *reinterpret_cast<const float4 *>(in + (long long)i * ld_in + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int t = 0; t < 16; t++) {
            const float4 ca = *reinterpret_cast<const float4 *>(&s.cs[t][ii0]);      // zero beyond sw: fma(-0, 0, a) == a
            const float c[GR] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
            for (int q = 0; q < GR; q++) {
                acc[q].x = gj_elim(acc[q].x, c[q], us[t].x);
                acc[q].y = gj_elim(acc[q].y, c[q], us[t].y);
                acc[q].z = gj_elim(acc[q].z, c[q], us[t].z);
                acc[q].w = gj_elim(acc[q].w, c[q], us[t].w);
            }
        }
#pragma unroll
        for (int q = 0; q < GR; q++) {
            const int i = i0 + ii0 + q;
            if (i >= n) continue;  // warp-uniform
            if (i >= r0 && i < r0 + sw) acc[q] = *reinterpret_cast<const float4 *>(&s.xf[i - r0][4 * lane]);
            if (!own_cols) *reinterpret_cast<float4 *>(out + (long long)i * ld_out + 4 * lane) = acc[q];
        }
    }
    TRACE(blockIdx.x == 3, 36);
}

cudaError_t debug_trace(int on, long long *out128) {
    cudaError_t e = cudaMemcpyToSymbol(g_trace_on, &on, sizeof(int));
    if (e == cudaSuccess && out128) e = cudaMemcpyFromSymbol(out128, g_trace, sizeof(long long) * 128);
    if (e == cudaSuccess) e = debug_trace_pk(on, out128 ? out128 + 96 : nullptr);  // batched kernel: slots 96..103
    return e;
}

// ------------------------------------------------------------------------------------------ launchers
template <int W, int R, int TH>
static cudaError_t launch_subpanel_t(int ncta, const float *in, long long ld_in, float *out, long long ld_out, int n,
                                     int k0, int s0, int sw, float *CmT, long long ldc, int *piv, float *pv, int *info,
                                     cudaStream_t st) {
    const size_t smem = sizeof(SubSmem) + (size_t)W * R * TH * sizeof(float);
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(subpanel_kernel<W, R, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(subpanel_kernel<W, R, TH>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(TH);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncta;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, subpanel_kernel<W, R, TH>, in, ld_in, out, ld_out, n, k0, s0, sw, CmT, ldc, piv, pv, info);
}

// Set by the shim while the trailing update runs on the tensor cores (MATINV_FLAG_TF32X3): the update is then ~2x shorter,
// the panel chain is the critical path at every n, and the shapes tuned for "fast" win over the shapes tuned for "small".
static int g_panel_critical = 0;
void panel_set_critical(int on) { g_panel_critical = on; }

// MATINV_SUBPANEL_SHAPE = WxRxTH forces one instantiation of the sub-panel kernel for every n it can hold (16 CTAs x TH
// threads x R rows), so that the shapes the large orders select -- 16x4x512 for 16384 < n <= 32768, 8x8x512 (8-wide
// sub-panels) above -- can be checked bit for bit against the oracle at orders the oracle finishes in seconds
// (tests/test_gpu_parity.py::test_subpanel_shapes_bit_identical).  Read once per process; 0 = not forced.
static int forced_shape() {
    static int shape = -1;
    if (shape < 0) {
        const char *e = getenv("MATINV_SUBPANEL_SHAPE");
        shape = 0;
        if (e && e[0]) {
            int w = 0, r = 0, th = 0;
            if (sscanf(e, "%dx%dx%d", &w, &r, &th) == 3) {
                if (w == 16 && r == 2 && th == 512) shape = 1;
                else if (w == 16 && r == 4 && th == 256) shape = 2;
                else if (w == 16 && r == 4 && th == 512) shape = 3;
                else if (w == 8 && r == 8 && th == 512) shape = 4;
                else if (w == 16 && r == 1 && th == 256) shape = 5;
                else if (w == 16 && r == 1 && th == 512) shape = 6;
            }
        }
    }
    return shape;
}
static int forced_capacity(int shape) { return shape == 1 ? 16384 : shape == 2 ? 16384 : shape == 3 ? 32768 : shape == 4 ? 65536 : shape == 5 ? 4096 : shape == 6 ? 8192 : 0; }

int subpanel_width(int n) {
    const int f = forced_shape();
    if (f && n <= forced_capacity(f)) return f == 4 ? 8 : 16;
    return (n > 32768) ? 8 : 16;
}
bool subpanel_supported(int n) { return n <= 65536; }

// Rows per cluster = ncta x TH x R.  512-thread CTAs with 2 rows per thread measured fastest at N=16384 on B200
// (49k cycles per launch; 256 threads x 4 rows: 62k).
cudaError_t launch_subpanel(const float *in, long long ld_in, float *out, long long ld_out, int n, int k0, int s0,
                            int sw, float *CmT, long long ldc, int *piv, float *pv, int *info, cudaStream_t st) {
#define SP_ARGS in, ld_in, out, ld_out, n, k0, s0, sw, CmT, ldc, piv, pv, info, st
    if (const int f = forced_shape(); f && n <= forced_capacity(f)) {
        // smallest power-of-two cluster that holds n rows (MATINV_SUBPANEL_CTAS overrides, e.g. 16 = the production size)
        static int forced_ctas = -1;
        if (forced_ctas < 0) {
            const char *e = getenv("MATINV_SUBPANEL_CTAS");
            forced_ctas = e ? atoi(e) : 0;
        }
        const int rows_per_cta = (f == 1) ? 1024 : (f == 2) ? 1024 : (f == 3) ? 2048 : (f == 5) ? 256 : (f == 6) ? 512 : 4096;
        int ncta = 1;
        while (ncta * rows_per_cta < n) ncta *= 2;
        if (forced_ctas > 0 && forced_ctas >= ncta && forced_ctas <= 16) ncta = forced_ctas;
        if (f == 1) return launch_subpanel_t<16, 2, 512>(ncta, SP_ARGS);
        if (f == 2) return launch_subpanel_t<16, 4, 256>(ncta, SP_ARGS);
        if (f == 3) return launch_subpanel_t<16, 4, 512>(ncta, SP_ARGS);
        if (f == 5) return launch_subpanel_t<16, 1, 256>(ncta, SP_ARGS);
        if (f == 6) return launch_subpanel_t<16, 1, 512>(ncta, SP_ARGS);
        return launch_subpanel_t<8, 8, 512>(ncta, SP_ARGS);
    }
    if (n <= 4096) {
        // more, smaller CTAs: the per-step chain (warp redux -> CTA arg max -> mailbox exchange) shortens with the CTA, and at
        // these orders the panel chain is the critical path (N=4096: 12.10 -> 11.67 ms, N=2048: 5.43 -> 5.18 ms;
        // profiles/r02_subpanel_shapes_small_n.txt)
        int ncta = 1;
        while (ncta * 256 < n) ncta *= 2;
        return launch_subpanel_t<16, 1, 256>(ncta, SP_ARGS);
    }
    if (n <= 8192) {
        int ncta = 1;
        while (ncta * 512 < n) ncta *= 2;
        return launch_subpanel_t<16, 1, 512>(ncta, SP_ARGS);
    }
    if (n <= 16384) {
        // Two shapes for the same 16 x n sub-panel.  512 threads x 2 rows is the faster kernel (25 vs 31 us) and is used
        // while the panel is on the critical path (measured better up to n = 12288: 80.6 vs 88.9 ms); near n = 16384 the panel
        // is hidden behind the trailing update and what
        // counts is what its CTAs displace: 256 threads x 4 rows stay within 128 registers, so a CTA fits beside one CTA of
        // the trailing update instead of taking the whole SM (N=16384: 171.8 -> 169.0 ms).  MATINV_K1_THREADS overrides.
        static int forced = -1;
        if (forced < 0) {
            const char *e = getenv("MATINV_K1_THREADS");
            forced = e ? atoi(e) : 0;
        }
        const int th = forced ? forced : ((n >= 15360 && !g_panel_critical) ? 256 : 512);
        if (th == 256) return launch_subpanel_t<16, 4, 256>(16, SP_ARGS);
        return launch_subpanel_t<16, 2, 512>(16, SP_ARGS);
    }
    if (n <= 32768) return launch_subpanel_t<16, 4, 512>(16, SP_ARGS);
    return launch_subpanel_t<8, 8, 512>(16, SP_ARGS);
#undef SP_ARGS
}

template <int ROWS>
static void launch_panel_update_t(const float *in, long long ld_in, float *out, long long ld_out, int n, int k0, int s0, int sw,
                                  int wfull, float *CmT, long long ldc, const int *piv, const float *pv, PanelState *ps, int kb,
                                  cudaStream_t st) {
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(panel_update_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpdSmem<ROWS>));
    }
    const int nblk = (n + ROWS - 1) / ROWS + 1;
    panel_update_kernel<ROWS><<<nblk, 256, sizeof(UpdSmem<ROWS>), st>>>(in, ld_in, out, ld_out, n, k0, s0, sw, wfull, CmT, ldc, piv,
                                                                        pv, ps, kb);
}

// Rows per CTA: 64 while the panel is on the critical path (12.3 ms against 14.2 ms at n = 4096, 57.3 against 58.1 at
// n = 10240), 256 above (80.6 against 81.5 ms at n = 12288) (it runs beside the trailing update and the time its CTAs
// hold SM slots is what counts: 171.7 ms against 178.5 ms at n = 16384).  MATINV_UPDATE_ROWS = 64 | 128 | 256 | 512 overrides.
void launch_panel_update(const float *in, long long ld_in, float *out, long long ld_out, int n, int k0, int s0, int sw,
                         int wfull, float *CmT, long long ldc, const int *piv, const float *pv, PanelState *ps, int kb,
                         cudaStream_t st) {
    static int forced = -1;
    if (forced < 0) {
        const char *e = getenv("MATINV_UPDATE_ROWS");
        forced = e ? atoi(e) : 0;
    }
    const int rows = forced ? forced : ((n <= 11264 || g_panel_critical) ? 64 : 256);
    if (rows == 64) launch_panel_update_t<64>(in, ld_in, out, ld_out, n, k0, s0, sw, wfull, CmT, ldc, piv, pv, ps, kb, st);
    else if (rows == 128) launch_panel_update_t<128>(in, ld_in, out, ld_out, n, k0, s0, sw, wfull, CmT, ldc, piv, pv, ps, kb, st);
    else if (rows == 512) launch_panel_update_t<512>(in, ld_in, out, ld_out, n, k0, s0, sw, wfull, CmT, ldc, piv, pv, ps, kb, st);
    else launch_panel_update_t<256>(in, ld_in, out, ld_out, n, k0, s0, sw, wfull, CmT, ldc, piv, pv, ps, kb, st);
}

// Whole panel (kb <= 128 columns starting at global row/column k0) of a column view `Wv` (leading dimension ld): 8
// x (sub-panel factor + in-panel update), ping-ponging through P0/P1.  piv is indexed by GLOBAL row (callers may pass
// a pointer biased by -k0).  Returns the number of kernels launched.
int launch_panel_factor(float *Wv, long long ld, int n, int k0, int kb, float *CmT, long long ldc, int *piv, float *pv,
                        int *info, PanelState *ps, float *P0, float *P1, cudaStream_t st) {
    float *P[2] = {P0, P1};
    const int sub = subpanel_width(n);
    const int ns = (kb + sub - 1) / sub;
    for (int s = 0; s < ns; s++) {
        const int s0 = s * sub;
        const int sw = (kb - s0 < sub) ? kb - s0 : sub;
        const float *in = (s == 0) ? Wv : P[(s - 1) & 1];
        const long long ld_in = (s == 0) ? ld : MATINV_NB;
        const bool to_w = (s == ns - 1) && ns > 1;
        float *out = to_w ? Wv : P[s & 1];
        const long long ld_out = to_w ? ld : MATINV_NB;
        launch_subpanel(in, ld_in, out, ld_out, n, k0, s0, sw, CmT, ldc, piv, pv, info, st);
        launch_panel_update(in, ld_in, out, ld_out, n, k0, s0, sw, sub, CmT, ldc, piv, pv, ps, kb, st);
    }
    if (ns == 1)  // single sub-panel: in == out would alias in the update kernel, so it went through P[0]
        cudaMemcpy2DAsync(Wv, ld * sizeof(float), P[0], MATINV_NB * sizeof(float), MATINV_NB * sizeof(float), n,
                          cudaMemcpyDeviceToDevice, st);
    return 2 * ns;
}
