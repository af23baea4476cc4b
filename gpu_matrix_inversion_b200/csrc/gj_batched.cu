// Batched small-n Gauss-Jordan: one CTA per matrix, matrix resident in shared memory for the whole
// inversion (north_star kernel 4).  The reference has no batched entry -- this is
// `for b: matrix_inv_32(A[b], n)` (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:11-395)
// collapsed into one launch; per matrix the arithmetic is the in-place form of SURVEY.md
// Appendix A.3, so results are bit-identical to the large-n path and to the oracle.
//
// v0 layout: a[n][n+1] floats in shared memory, 256 threads, three barriers per column.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

__global__ void __launch_bounds__(256)
batched_smem_kernel(const float *__restrict__ A, int n, long long batch, float *__restrict__ X,
                    int *__restrict__ info) {
    extern __shared__ float sm[];
    const int ldm = n + 1;
    float *a = sm;                        // n * (n+1)
    float *urow = a + n * ldm;            // n
    float *ccol = urow + n;               // n
    int *piv = reinterpret_cast<int *>(ccol + n);  // n
    __shared__ u64 skey[8];
    __shared__ int sinfo;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const float *Ab = A + b * (long long)n * n;
        float *Xb = X + b * (long long)n * n;
        for (int e = tid; e < n * n; e += 256) a[(e / n) * ldm + (e % n)] = Ab[e];
        if (tid == 0) sinfo = 0;
        __syncthreads();

        for (int r = 0; r < n; r++) {
            // (1) pivot search: rows r..n-1 of column r
            u64 k = 0;
            for (int i = r + tid; i < n; i += 256) {
                const u64 kk = gj_key(a[i * ldm + r], i, i == r);
                k = kk > k ? kk : k;
            }
            k = warp_max_u64(k);
            if (lane == 0) skey[warp] = k;
            __syncthreads();
            u64 best = skey[0];
#pragma unroll
            for (int w = 1; w < 8; w++) best = skey[w] > best ? skey[w] : best;
            const int p = gj_key_row(best);
            const float v = gj_key_value(best);
            if (tid == 0) {
                piv[r] = p;
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
            }
            // (2) swap + normalise: read phase
            float rr = 0.f, pp = 0.f, cc = 0.f;
            if (tid < n) {
                rr = a[r * ldm + tid];
                pp = a[p * ldm + tid];
                cc = a[tid * ldm + r];
            }
            const float arr = a[r * ldm + r];
            __syncthreads();
            if (tid < n) {
                const float u = (tid == r) ? 1.0f / v : pp / v;
                urow[tid] = u;
                a[r * ldm + tid] = u;
                if (p != r) a[p * ldm + tid] = rr;
                ccol[tid] = (tid == r) ? 0.0f : ((tid == p) ? arr : cc);
            }
            __syncthreads();
            // (3) rank-1 update of every row but r
            for (int e = tid; e < n * n; e += 256) {
                const int i = e / n, j = e - i * n;
                if (i == r) continue;
                const float c = ccol[i], u = urow[j];
                a[i * ldm + j] = (j == r) ? fmaf(-c, u, 0.0f) : gj_elim(a[i * ldm + j], c, u);
            }
            __syncthreads();
        }
        // deferred column permutation (see gj_finish.cu:colperm_kernel)
        int *colsrc = reinterpret_cast<int *>(urow);
        __syncthreads();
        if (tid < n) {
            int c = tid;
            for (int r = 0; r < n; r++) {
                if (c < r) break;
                const int p = piv[r];
                if (c == r) c = p;
                else if (c == p) c = r;
            }
            colsrc[tid] = c;
        }
        __syncthreads();
        bool bad = false;
        for (int e = tid; e < n * n; e += 256) {
            const int i = e / n, j = e - i * n;
            const float x = a[i * ldm + colsrc[j]];
            bad |= !isfinite(x);
            Xb[e] = x;
        }
        const int anybad = __syncthreads_or(bad);
        if (tid == 0 && info) info[b] = sinfo ? sinfo : (anybad ? -1 : 0);
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// v1: ONE WARP PER MATRIX, matrix resident in REGISTERS (n = 64: lane l owns physical rows l and
// l+32, 128 registers; n = 32: one row per lane).
//   * row interchanges are implicit (each row carries its logical position, see gj_subpanel.cu)
//   * the column window rotates left by one per step, so the step loop is rolled and every register
//     index is static: the current column is always a[.][0], the new inverse column enters at N-1
//   * pivot search: two redux.sync; pivot row broadcast and the distributed true division go through
//     2 x N floats of shared memory; the rank-1 update is N FFMA per owned row
//   * the deferred column permutation is kept incrementally (qinv) and applied when the result is
//     staged through shared memory for coalesced stores
template <int N>
__global__ void __launch_bounds__(128, 2)
batched_reg_kernel(const float *__restrict__ A, long long batch, float *__restrict__ X, int *__restrict__ info) {
    constexpr int RS = N / 32;
    constexpr int LD = N + 1;
    constexpr int PER_WARP = ((N * LD + N + 3) / 4) * 4;  // floats; keeps every warp's base 16-byte aligned
    extern __shared__ __align__(16) float smem_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *stage = smem_f + warp * PER_WARP;
    float *raw = stage, *urot = stage + N;
    int *qinv = reinterpret_cast<int *>(stage + N * LD);
    const unsigned urot_s = (unsigned)__cvta_generic_to_shared(urot);

    for (long long b = (long long)blockIdx.x * 4 + warp; b < batch; b += (long long)gridDim.x * 4) {
        const float *Ab = A + b * (long long)(N * N);
        float a[RS][N];
        int lpos[RS];
        bool used[RS];
#pragma unroll
        for (int s = 0; s < RS; s++) {
            const float4 *src = reinterpret_cast<const float4 *>(Ab + (s * 32 + lane) * N);
#pragma unroll
            for (int f = 0; f < N / 4; f++) {
                const float4 v4 = src[f];
                a[s][4 * f] = v4.x; a[s][4 * f + 1] = v4.y; a[s][4 * f + 2] = v4.z; a[s][4 * f + 3] = v4.w;
            }
            lpos[s] = s * 32 + lane;
            used[s] = false;
            qinv[s * 32 + lane] = s * 32 + lane;
        }
        int sinfo = 0;
        __syncwarp();

        // Steps run in groups of 4 with STATIC column indices inside a group; after each group the register
        // window rotates left by 4 (a[.][j] <- a[.][j+4]), so the loop over groups is rolled (small code) at the
        // cost of N/4 register moves per step and row.  Window position w of group g is column 4g + w (mod N).
#pragma unroll 1
        for (int g = 0; g < N / 4; g++) {
#pragma unroll
            for (int tc = 0; tc < 4; tc++) {
                const int r = 4 * g + tc;
                // ---- (1) pivot search over the not-yet-used rows (== logical positions >= r)
                unsigned mag = 0;
                int cand = 0x7FFFFFFF;
                bool has = false;
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    if (!used[s]) {
                        const unsigned mq = gj_mag(a[s][tc], lpos[s] == r);
                        if (!has || mq > mag || (mq == mag && lpos[s] < cand)) { mag = mq; cand = lpos[s]; has = true; }
                    }
                }
                const unsigned gm = __reduce_max_sync(0xffffffffu, has ? mag : 0u);
                const int p = (int)__reduce_min_sync(0xffffffffu, (has && mag == gm) ? (unsigned)cand : 0x7FFFFFFFu);
                bool own[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) own[s] = !used[s] && lpos[s] == p;
                const unsigned b0 = __ballot_sync(0xffffffffu, own[0]);
                unsigned b1 = 0;
                if (RS > 1) b1 = __ballot_sync(0xffffffffu, own[RS - 1]);
                const int Sp = b0 ? 0 : 1;               // warp-uniform slot of the pivot row
                const int Lp = __ffs(b0 ? b0 : b1) - 1;  // its lane
                const float v = __shfl_sync(0xffffffffu, (Sp == 0) ? a[0][tc] : a[RS - 1][tc], Lp);
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
                // ---- (2) publish the raw pivot row (window order)
                if (lane == Lp) {
                    if (Sp == 0) {
#pragma unroll
                        for (int f = 0; f < N / 4; f++)
                            reinterpret_cast<float4 *>(raw)[f] = make_float4(a[0][4 * f], a[0][4 * f + 1], a[0][4 * f + 2], a[0][4 * f + 3]);
                    } else {
#pragma unroll
                        for (int f = 0; f < N / 4; f++)
                            reinterpret_cast<float4 *>(raw)[f] =
                                make_float4(a[RS - 1][4 * f], a[RS - 1][4 * f + 1], a[RS - 1][4 * f + 2], a[RS - 1][4 * f + 3]);
                    }
                }
                __syncwarp();
                // ---- (3) true division, RS elements per lane; the pivot position receives 1/v
#pragma unroll
                for (int k = 0; k < RS; k++) {
                    const int pos = lane * RS + k;
                    const float val = raw[pos];
                    urot[pos] = (pos == tc) ? 1.0f / v : val / v;
                }
                __syncwarp();
                // ---- (4) rank-1 update: every owned row (the pivot row too -- its result is thrown away) takes the
                //          FMA pass with u consumed straight from shared memory, one LDS.128 per four columns shared by
                //          all rows of the lane; then the pivot lane reloads its pivot row as u (16 LDS.128, cheaper
                //          than N predicated moves)
                float cm[RS];
#pragma unroll
                for (int s = 0; s < RS; s++) cm[s] = a[s][tc];
#pragma unroll
                for (int f = 0; f < N / 4; f++) {
                    const float4 u4 = reinterpret_cast<const float4 *>(urot)[f];
                    const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                    for (int s = 0; s < RS; s++)
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = 4 * f + k;
                            a[s][j] = (j == tc) ? fmaf(-cm[s], uu[k], 0.0f) : gj_elim(a[s][j], cm[s], uu[k]);
                        }
                }
                if (lane == Lp) {
#pragma unroll
                    for (int s = 0; s < RS; s++)
                        if (s == Sp) {
#pragma unroll
                            for (int f = 0; f < N / 4; f++)
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(a[s][4 * f]), "=f"(a[s][4 * f + 1]), "=f"(a[s][4 * f + 2]), "=f"(a[s][4 * f + 3])
                                             : "r"(urot_s + 16 * f));
                        }
                }
                // ---- (5) bookkeeping: logical positions, column permutation
#pragma unroll
                for (int s = 0; s < RS; s++) {
                    if (own[s]) { used[s] = true; lpos[s] = r; }
                    else if (!used[s] && lpos[s] == r) lpos[s] = p;
                }
                if (lane == 0) { const int q1 = qinv[r], q2 = qinv[p]; qinv[r] = q2; qinv[p] = q1; }
                __syncwarp();
            }
            // rotate the window left by 4
#pragma unroll
            for (int s = 0; s < RS; s++) {
                const float t0 = a[s][0], t1 = a[s][1], t2 = a[s][2], t3 = a[s][3];
#pragma unroll
                for (int j = 0; j < N - 4; j++) a[s][j] = a[s][j + 4];
                a[s][N - 4] = t0; a[s][N - 3] = t1; a[s][N - 2] = t2; a[s][N - 1] = t3;
            }
        }

        // ---- result: X[lpos][qinv[c]] = a[.][c], staged through shared memory for coalesced stores
#pragma unroll
        for (int s = 0; s < RS; s++) {
            float *row = stage + lpos[s] * LD;
#pragma unroll
            for (int c = 0; c < N; c++) row[qinv[c]] = a[s][c];
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

// ------------------------------------------------------------------------------------------------
// v2: TWO WARPS PER 64x64 MATRIX, one row per thread (64 data registers).  Same arithmetic and the same tricks as
// v1 (implicit row interchanges, groups of 4 static steps + window rotation, incremental column permutation), but
// a third of the registers: 5 CTAs x 128 threads per SM = 20 warps hide the per-step dependency chain
// (arg max -> publish pivot row -> distributed division -> update) that left v1 latency-bound at 8 warps per SM.
// The two warps of a matrix meet at a 64-thread named barrier three times per step.
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(128, 5)
batched_row64_kernel(const float *__restrict__ A, long long batch, float *__restrict__ X, int *__restrict__ info) {
    constexpr int N = 64, LD = N + 1;
    constexpr int PER_PAIR = ((N * LD + N + 4 + 3) / 4) * 4;  // floats: stage (aliases raw,u) + qinv + keys
    extern __shared__ __align__(16) float smem_f[];
    const int pair = threadIdx.x >> 6, t64 = threadIdx.x & 63, wp = t64 >> 5, lane = threadIdx.x & 31;
    float *stage = smem_f + pair * PER_PAIR;
    float *raw = stage, *ub = stage + N;
    int *qinv = reinterpret_cast<int *>(stage + N * LD);
    u64 *keys = reinterpret_cast<u64 *>(stage + N * LD + N);   // [2], 8-byte aligned (N*LD+N = 4224 floats)
    const unsigned ub_s = (unsigned)__cvta_generic_to_shared(ub);
    const int bar = 1 + pair;

    for (long long b = (long long)blockIdx.x * 2 + pair; b < batch; b += (long long)gridDim.x * 2) {
        const float *Ab = A + b * (long long)(N * N);
        float a[N];
        {
            const float4 *src = reinterpret_cast<const float4 *>(Ab + t64 * N);
#pragma unroll
            for (int f = 0; f < N / 4; f++) {
                const float4 v4 = src[f];
                a[4 * f] = v4.x; a[4 * f + 1] = v4.y; a[4 * f + 2] = v4.z; a[4 * f + 3] = v4.w;
            }
        }
        int lpos = t64;
        bool used = false;
        int sinfo = 0;
        qinv[t64] = t64;
        pair_barrier(bar);

#pragma unroll 1
        for (int g = 0; g < N / 4; g++) {
#pragma unroll
            for (int tc = 0; tc < 4; tc++) {
                const int r = 4 * g + tc;
                // ---- (1) arg max over the unused rows: warp redux, then the two warps exchange their keys
                const unsigned mag = used ? 0u : gj_mag(a[tc], lpos == r);
                const unsigned gm = __reduce_max_sync(0xffffffffu, mag);
                const unsigned pw = __reduce_min_sync(0xffffffffu, (!used && mag == gm) ? (unsigned)lpos : 0x7FFFFFFFu);
                if (pw == 0x7FFFFFFFu) {
                    if (lane == 0) keys[wp] = 0;
                } else if (!used && mag == gm && (unsigned)lpos == pw) {
                    keys[wp] = gj_key_from(mag, lpos, a[tc]);
                }
                pair_barrier(bar);
                const u64 k0 = keys[0], k1 = keys[1];
                const u64 kb = k0 > k1 ? k0 : k1;
                const int p = gj_key_row(kb);
                const float v = gj_key_value(kb);
                if (gj_bad_pivot(v) && sinfo == 0) sinfo = r + 1;
                const bool own = !used && lpos == p;
                // ---- (2) the owner publishes its raw row (window order)
                if (own) {
#pragma unroll
                    for (int f = 0; f < N / 4; f++)
                        reinterpret_cast<float4 *>(raw)[f] = make_float4(a[4 * f], a[4 * f + 1], a[4 * f + 2], a[4 * f + 3]);
                }
                pair_barrier(bar);
                // ---- (3) true division, one element per thread; the pivot position receives 1/v
                ub[t64] = (t64 == tc) ? 1.0f / v : raw[t64] / v;
                pair_barrier(bar);
                // ---- (4) rank-1 update, u consumed four values at a time
                if (own) {
#pragma unroll
                    for (int f = 0; f < N / 4; f++)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(a[4 * f]), "=f"(a[4 * f + 1]), "=f"(a[4 * f + 2]), "=f"(a[4 * f + 3])
                                     : "r"(ub_s + 16 * f));
                    used = true;
                    lpos = r;
                } else {
                    const float c = a[tc];
#pragma unroll
                    for (int f = 0; f < N / 4; f++) {
                        const float4 u4 = reinterpret_cast<const float4 *>(ub)[f];
                        const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = 4 * f + k;
                            a[j] = (j == tc) ? fmaf(-c, uu[k], 0.0f) : gj_elim(a[j], c, uu[k]);
                        }
                    }
                    if (!used && lpos == r) lpos = p;
                }
                if (t64 == 0) { const int q1 = qinv[r], q2 = qinv[p]; qinv[r] = q2; qinv[p] = q1; }
            }
            // rotate the window left by 4
            {
                const float t0 = a[0], t1 = a[1], t2 = a[2], t3 = a[3];
#pragma unroll
                for (int j = 0; j < N - 4; j++) a[j] = a[j + 4];
                a[N - 4] = t0; a[N - 3] = t1; a[N - 2] = t2; a[N - 1] = t3;
            }
        }

        // ---- result: X[lpos][qinv[c]] = a[c], staged through shared memory for coalesced stores
        pair_barrier(bar);
        {
            float *row = stage + lpos * LD;
#pragma unroll
            for (int c = 0; c < N; c++) row[qinv[c]] = a[c];
        }
        pair_barrier(bar);
        float *Xb = X + b * (long long)(N * N);
        bool bad = false;
#pragma unroll 4
        for (int i = wp; i < N; i += 2) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const float xv = stage[i * LD + lane + 32 * k];
                bad |= !isfinite(xv);
                Xb[i * N + lane + 32 * k] = xv;
            }
        }
        const int anybad = __any_sync(0xffffffffu, bad);
        if (lane == 0) keys[wp] = anybad;
        pair_barrier(bar);
        if (t64 == 0 && info) info[b] = sinfo ? sinfo : ((keys[0] | keys[1]) ? -1 : 0);
        pair_barrier(bar);
    }
}

static cudaError_t launch_batched_row64(const float *A, long long batch, float *X, int *info, cudaStream_t st) {
    constexpr int N = 64;
    constexpr int PER_PAIR = ((N * (N + 1) + N + 4 + 3) / 4) * 4;
    const size_t smem = 2 * PER_PAIR * sizeof(float);
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(batched_row64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    long long grid = (batch + 1) / 2;
    const long long cap = 148ll * 5 * 16;
    if (grid > cap) grid = cap;
    batched_row64_kernel<<<(unsigned)grid, 128, smem, st>>>(A, batch, X, info);
    return cudaGetLastError();
}

template <int N>
static cudaError_t launch_batched_reg(const float *A, long long batch, float *X, int *info, cudaStream_t st) {
    constexpr int PER_WARP = ((N * (N + 1) + N + 3) / 4) * 4;
    const size_t smem = 4 * PER_WARP * sizeof(float);
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(batched_reg_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    long long grid = (batch + 3) / 4;
    const long long cap = 148ll * 2 * 16;
    if (grid > cap) grid = cap;
    batched_reg_kernel<N><<<(unsigned)grid, 128, smem, st>>>(A, batch, X, info);
    return cudaGetLastError();
}

// MATINV_BATCHED: 0 = shared-memory kernel (v0), 1 = one warp per matrix, scalar FMAs (v1), 2 = two warps per matrix (v2),
// 3 = one warp per matrix on packed pairs, gj_batched_pk.cu (v3, default: 2.50e7 inversions/s on B200 against 2.19e7 for v1 and
// 1.7e7 for v2), 4 = v3's layout blocked in 8-column panels with look-ahead, gj_batched_blk.cu (v4: same rate, see its header)
static int batched_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("MATINV_BATCHED");
        mode = e ? atoi(e) : 3;
    }
    return mode;
}

cudaError_t launch_batched(const float *A, int n, long long batch, float *X, int *info, cudaStream_t st) {
    if (batched_mode() >= 1) {
        if ((n == 64 || n == 32) && batched_mode() >= 4) return launch_batched_blk(A, n, batch, X, info, st);
        if ((n == 64 || n == 32) && batched_mode() == 3) return launch_batched_pk(A, n, batch, X, info, st);
        if (n == 64 && batched_mode() == 2) return launch_batched_row64(A, batch, X, info, st);
        if (n == 64) return launch_batched_reg<64>(A, batch, X, info, st);
        if (n == 32) return launch_batched_reg<32>(A, batch, X, info, st);
    }
    const size_t smem = ((size_t)n * (n + 1) + 3 * (size_t)n) * sizeof(float);
    static bool configured[64] = {};
    if (first_use_on_device(configured)) {
        cudaFuncSetAttribute(batched_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 129 * 4 + 3 * 128 * 4);
    }
    long long grid = batch;
    if (grid > (1ll << 20)) grid = 1ll << 20;
    batched_smem_kernel<<<(unsigned)grid, 256, smem, st>>>(A, n, batch, X, info);
    return cudaGetLastError();
}
