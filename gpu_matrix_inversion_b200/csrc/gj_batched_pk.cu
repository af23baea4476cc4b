// Batched small-n Gauss-Jordan, v3: one warp per matrix, matrix resident in registers as PACKED PAIRS.
//
// Same arithmetic, layout and bookkeeping as v1 (gj_batched.cu: lane l owns physical rows l and l+32, implicit row
// interchanges through logical positions, static column indices inside a group of G steps and a window rotation
// between groups), so results stay bit-identical to the oracle.  What changes is everything around the FMAs -- ncu on
// v1 shows 360 warp instructions per pivot step for 128 FFMA and an issue slot busy 42 % of the time with 8 warps
// per SM, the rest being dependent scalar code whose latencies nothing hides:
//   * the rank-1 update is N/2 fma.rn.f32x2 per row (FFMA2: the same IEEE fma on both halves); the pivot column is
//     zeroed first so that it takes the same FMA as every other column (fma(-c, 1/v, +0) is the in-place inverse entry)
//   * no `used` flags (a row is spent iff its logical position is < r) and no ballots / shuffles: the slot of the pivot
//     row rides in the low bit of the redux.min operand, the pivot value is read back from the published raw row
//   * the owner publishes and re-loads the pivot row with predicated 128-bit shared-memory accesses straight from /
//     into the row's registers, under a warp-uniform branch on the slot (no divergent region)
//   * one true division per element of the pivot row (numerator 1 at the pivot position), two __syncwarp per step;
//     lane 0 keeps the column permutation off the critical path
//
// What bounds it (tools/trace_batched.py, tools/sass_stalls.py): a pivot step is one dependency chain of about 1000
// cycles for a warp alone on its scheduler -- search (two redux) 190, publish 175, division 275, update 280, reload
// 135 -- against 290 issued instructions, and the register file admits two such warps per scheduler.  Tried and
// measured without gain: issuing the next step's search from inside the update (the chain only moves), a division
// with the reciprocal shared between the two numerators, groups of 8 steps, 3 warps per scheduler at 168 registers.
#include "common.cuh"
#include "kernels.h"

__device__ __forceinline__ u64 pk_pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void pk_unpack(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 pk_fma(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// One row of packed pairs <-> shared memory, 128 bits at a time, predicated: the caller branches on the (warp-uniform)
// slot of the pivot row, every lane issues the access and only the owner performs it -- no divergent region.  The slot
// number inside the asm text keeps the two arms textually distinct (identical asm statements in sibling branches get
// merged and the data then travels through temporaries and N moves); the reload is volatile because ptxas otherwise
// recognises the address of the u loads of the update and replaces the load by 64 predicated moves.
template <int S, int NP>
__device__ __forceinline__ void row_publish_if(unsigned addr, const u64 (&row)[NP], bool pr) {
#pragma unroll
    for (int f = 0; f < NP / 2; f++)
        asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.b64 [%0], {%1, %2}; }  // slot %4" ::"r"(addr + 16 * f),
                     "l"(row[2 * f]), "l"(row[2 * f + 1]), "r"((int)pr), "n"(S) : "memory");
}
template <int S, int NP>
__device__ __forceinline__ void row_reload_if(unsigned addr, u64 (&row)[NP], bool pr) {
#pragma unroll
    for (int f = 0; f < NP / 2; f++)
        asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q ld.volatile.shared.v2.b64 {%0, %1}, [%2]; }  // slot %4"
                     : "+l"(row[2 * f]), "+l"(row[2 * f + 1]) : "r"(addr + 16 * f), "r"((int)pr), "n"(S) : "memory");
}

// Optional per-phase cycle accounting of warp 0 of CTA 0 (matinv_debug_trace slots 96..103), TRACE builds only.
__device__ long long g_pk_trace[8];
static int g_pk_trace_on = 0;
#define PK_MARK(i) do { if (TRACE) { const unsigned tn = clock(); tacc[i] += tn - tprev; tprev = tn; } } while (0)

template <int N, int G, int WPC, int CPS, bool TRACE>
__global__ void __launch_bounds__(32 * WPC, CPS)
batched_pk_kernel(const float *__restrict__ A, long long batch, float *__restrict__ X, int *__restrict__ info) {
    static_assert(N % G == 0 && G % 2 == 0, "group size");
    constexpr int RS = N / 32;
    constexpr int LD = N + 1;
    constexpr int PER_WARP = ((N * LD + N + 3) / 4) * 4;  // floats; keeps every warp's base 16-byte aligned
    extern __shared__ __align__(16) float smem_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *stage = smem_f + warp * PER_WARP;
    float *raw = stage, *urot = stage + N;
    int *qinv = reinterpret_cast<int *>(stage + N * LD);
    const unsigned raw_s = (unsigned)__cvta_generic_to_shared(raw);
    const unsigned urot_s = (unsigned)__cvta_generic_to_shared(urot);

    for (long long b = (long long)blockIdx.x * WPC + warp; b < batch; b += (long long)gridDim.x * WPC) {
        const float *Ab = A + b * (long long)(N * N);
        u64 a2[RS][N / 2];
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
        int sinfo = 0;
        __syncwarp();
        unsigned tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned tprev = TRACE ? clock() : 0u;

#pragma unroll 1
        for (int g = 0; g < N / G; g++) {
#pragma unroll
            for (int tc = 0; tc < G; tc++) {
                const int r = G * g + tc;
                // ---- (1) arg max over the unspent rows (logical position >= r), lowest position on ties
                float cm[RS];
                unsigned mag = 0;
                unsigned cand = 0x7FFFFFFFu;
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    float lo, hi;
                    pk_unpack(a2[s][tc >> 1], lo, hi);
                    cm[s] = (tc & 1) ? hi : lo;
                    const bool live = lpos[s] >= r;
                    const unsigned mq = live ? gj_mag(cm[s], lpos[s] == r) : 0u;
                    const unsigned lq = live ? (((unsigned)lpos[s] << 1) | (unsigned)s) : 0x7FFFFFFFu;  // slot rides along
                    if (s == 0 || mq > mag || (mq == mag && lq < cand)) { mag = mq; cand = lq; }
                }
                const unsigned gm = __reduce_max_sync(0xffffffffu, mag);
                const unsigned pe = __reduce_min_sync(0xffffffffu, mag == gm ? cand : 0x7FFFFFFFu);
                const int p = (int)(pe >> 1);
                const int Sp = (int)(pe & 1u);  // warp-uniform slot of the pivot row
                bool own[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) own[s] = lpos[s] == p;
                PK_MARK(0);
                // ---- (2) the owner publishes the raw pivot row (window order)
                if (RS == 1 || Sp == 0) row_publish_if<0>(raw_s, a2[0], own[0]);
                else row_publish_if<1>(raw_s, a2[RS - 1], own[RS - 1]);
                __syncwarp();
                PK_MARK(1);
                // ---- (3) true division, RS elements per lane; the pivot position receives 1/v
                const float v = raw[tc];
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
#pragma unroll
                for (int k = 0; k < RS; k++) {
                    const int pos = lane + 32 * k;
                    const float num = (pos == tc) ? 1.0f : raw[pos];
                    urot[pos] = __fdiv_rn(num, v);
                }
                __syncwarp();
                PK_MARK(2);
                // ---- (4) rank-1 update on packed pairs; the pivot column starts from +0 so that it receives -c/v.
                //          Every owned row takes the pass (the pivot row's result is thrown away in (5)).
                u64 ncm[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    float lo, hi;
                    pk_unpack(a2[s][tc >> 1], lo, hi);
                    a2[s][tc >> 1] = (tc & 1) ? pk_pack(lo, 0.0f) : pk_pack(0.0f, hi);
                    ncm[s] = pk_pack(-cm[s], -cm[s]);
                }
#pragma unroll
                for (int f = 0; f < N / 4; f++) {
                    const ulonglong2 u4 = reinterpret_cast<const ulonglong2 *>(urot)[f];
#pragma unroll
                    for (int s = 0; s < RS; s++) {
                        a2[s][2 * f] = pk_fma(ncm[s], u4.x, a2[s][2 * f]);
                        a2[s][2 * f + 1] = pk_fma(ncm[s], u4.y, a2[s][2 * f + 1]);
                    }
                }
                PK_MARK(3);
                // ---- (5) the pivot row becomes u: the owner reloads it straight into the row's registers
                if (RS == 1 || Sp == 0) row_reload_if<0>(urot_s, a2[0], own[0]);
                else row_reload_if<1>(urot_s, a2[RS - 1], own[RS - 1]);
                PK_MARK(4);
                // ---- (6) bookkeeping: logical positions; lane 0 keeps the column permutation
#pragma unroll
                for (int s = 0; s < RS; s++) lpos[s] = own[s] ? r : (lpos[s] == r ? p : lpos[s]);
                if (lane == 0) {
                    const int q1 = qinv[r], q2 = qinv[p];
                    qinv[r] = q2;
                    qinv[p] = q1;
                }
            }
            PK_MARK(5);
            // rotate the window left by G columns
#pragma unroll
            for (int s = 0; s < RS; s++) {
                u64 t[G / 2];
#pragma unroll
                for (int j = 0; j < G / 2; j++) t[j] = a2[s][j];
#pragma unroll
                for (int j = 0; j < (N - G) / 2; j++) a2[s][j] = a2[s][j + G / 2];
#pragma unroll
                for (int j = 0; j < G / 2; j++) a2[s][(N - G) / 2 + j] = t[j];
            }
        }
        __syncwarp();
        PK_MARK(6);
        if (TRACE && blockIdx.x == 0 && threadIdx.x == 0 && b == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) g_pk_trace[i] = tacc[i];
        }

        // ---- result: X[lpos][qinv[c]] = a[.][c], staged through shared memory for coalesced stores
#pragma unroll
        for (int s = 0; s < RS; s++) {
            float *row = stage + lpos[s] * LD;
#pragma unroll
            for (int c = 0; c < N / 2; c++) {
                float lo, hi;
                pk_unpack(a2[s][c], lo, hi);
                row[qinv[2 * c]] = lo;
                row[qinv[2 * c + 1]] = hi;
            }
        }
        __syncwarp();
        float *Xb = X + b * (long long)(N * N);
        bool bad = false;
#pragma unroll 4
        for (int i = 0; i < N; i++) {
#pragma unroll
            for (int k = 0; k < RS; k++) {
                const float xv = stage[i * LD + lane + 32 * k];
                bad |= !isfinite(xv);
                Xb[i * N + lane + 32 * k] = xv;
            }
        }
        const int anybad = __any_sync(0xffffffffu, bad);
        if (lane == 0 && info) info[b] = sinfo ? sinfo : (anybad ? -1 : 0);
        __syncwarp();
    }
}

template <int N, int G, int WPC, int CPS>
static cudaError_t launch_pk(const float *A, long long batch, float *X, int *info, cudaStream_t st) {
    constexpr int PER_WARP = ((N * (N + 1) + N + 3) / 4) * 4;
    const size_t smem = WPC * PER_WARP * sizeof(float);
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(batched_pk_kernel<N, G, WPC, CPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(batched_pk_kernel<N, G, WPC, CPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    long long grid = (batch + WPC - 1) / WPC;
    const long long cap = 148ll * CPS * 16;
    if (grid > cap) grid = cap;
    if (g_pk_trace_on) batched_pk_kernel<N, G, WPC, CPS, true><<<(unsigned)grid, 32 * WPC, smem, st>>>(A, batch, X, info);
    else batched_pk_kernel<N, G, WPC, CPS, false><<<(unsigned)grid, 32 * WPC, smem, st>>>(A, batch, X, info);
    return cudaGetLastError();
}

// n in {32, 64}.  n = 64 runs 4 warps per CTA and 2 CTAs per SM: a scheduler's register file holds two warps of 255
// registers or three of 168, and at 168 the kernel spills half a dozen scalars into the per-step dependency chain
// (measured: 2.1e7 inversions/s with 12 warps per SM against 2.5e7 with 8).
cudaError_t launch_batched_pk(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st) {
    if (n == 64) return launch_pk<64, 4, 4, 2>(A, batch, X, info, st);
    if (n == 32) return launch_pk<32, 4, 4, 4>(A, batch, X, info, st);
    return cudaErrorInvalidValue;
}

// slots [0, 8) of out8: cycles warp 0 of CTA 0 spent in search / publish / division / update / reload / bookkeeping /
// rotation for its first matrix (sums over all steps)
cudaError_t debug_trace_pk(int on, long long *out8) {
    g_pk_trace_on = on;
    if (out8) return cudaMemcpyFromSymbol(out8, g_pk_trace, sizeof(long long) * 8);
    return cudaSuccess;
}
