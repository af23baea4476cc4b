/*
 * gj_oracle.c -- CPU restatement of the reference's Gauss-Jordan inversion.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (gpu_matrix_inversion_b200/csrc) has no CPU fallback.
 *
 * PARITY STATUS: pinned.  The reference ships no golden vectors / known-answer tests and
 * its OpenCL host needs an ICD that this image does not have, so the reference's own translation units are
 * compiled from /root/reference against oracle/minicl (a CPU OpenCL runtime written for this repo) and executed;
 * tests/test_oracle_vs_reference.py requires this restatement (with GJ_QUIRK, the as-written pivot search) to
 * reproduce those outputs bit for bit -- see oracle/README.md.
 *
 * Reference (paths under /root/reference, LIB = Matlab/mat_inv_32/mat_inv_32,
 * SOL = matrix_inv_solution/matrix_inversion_solution/matrix_inversion):
 *   LIB/mat_inv_32.cpp:177-192  makeAugmentedMatrix   M = [A | I]
 *   LIB/mat_inv_32.cpp:61-132   maxPivot/finalMaxPivot pivot = arg max |M[i][r]|, i >= r, strict '>'
 *   LIB/mat_inv_32.cpp:154-173  pivotElementsKernel   swap rows r, p  iff p != r
 *   LIB/mat_inv_32.cpp:138-150  fixRowKernel          M[r][:] /= pivot VALUE (true division)
 *   LIB/mat_inv_32.cpp:13-57    fixColumnKernel       M[i][:] -= M[i][r]*M[r][:]  (i != r, M[i][r] != 0)
 *   LIB/mat_inv_32.cpp:195-203  getInvertedMatrix     X = M[:, N:2N]
 *   LIB/mat_inv_32.cpp:317-362  step order            search -> swap -> scale -> eliminate
 *   SOL/matrix_inversion_FP32.cpp:814-835             left half != I exactly  ==> return {}
 *
 * Three algebraically different but (for finite data) BIT-IDENTICAL formulations are kept so
 * the tests can prove the invariants the CUDA kernels rely on:
 *   gj_aug_*      A.1/A.2  explicit N x 2N [A|I], 4N^3 flops     (the reference's data layout)
 *   gj_inplace_*  A.3      in-place N x N, deferred column permutation, 2N^3 flops
 *   gj_blocked_*  A.4      right-looking blocked form: panel -> swaps -> row-block recurrence
 *                          -> trailing update with a k-sequential FMA chain seeded from C
 *
 * Arithmetic: default a <- fmaf(-c, u, a) (what nvcc emits for a - c*u); GJ_NOFMA computes
 * a - fl(c*u) to bracket what an OpenCL compiler may have done.  Row scale is IEEE x / v.
 *
 * Return value ("info"): 0 = ok; r+1 = pivot at step r was 0 or non-finite (singular);
 * -1 = the inverse holds a non-finite entry (overflow) -- all non-zero values mean the
 * library returns an empty vector.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GJ_NOFMA 1 /* flags bit0: a - fl(c*u) instead of fmaf(-c,u,a) */
#define GJ_QUIRK 2 /* flags bit1 (gj_aug_* only): pivot search AS WRITTEN in the reference, see below */
#define GJ_NOPIVOT 4 /* flags bit2: pivot = diagonal entry, no interchange (matrix_inversion_no_pivots.cpp:41-76) */

/* The reference's pivot search exactly as its two kernels behave (LIB/mat_inv_32.cpp:61-132), for n % 256 == 0:
 * per 256-row work-group a tree reduction over the WRONG window [0, lim) of the local array with receivers
 * restricted to rows >= r, then a sequential strict-'>' scan of the per-group results.
 * Not the parity target of the product (north_star pins the intended arg max); it exists so that the whole
 * restatement -- swap, scale, eliminate, extract -- can be compared BIT FOR BIT with outputs of the unmodified
 * reference on inputs that need real row interchanges (tests/test_oracle_vs_reference.py). */
#define DEFINE_QUIRK(T, SUF, ABS)                                                                          \
    static int quirk_pivot_##SUF(const T *M, size_t ld, int n, int r, T *value) {                          \
        T bestv = 0, besti = 0;                                                                            \
        const int ngroups = n / 256;                                                                       \
        for (int g = 0; g < ngroups; g++) {                                                                \
            T ox = 0, oy = 0;                                                                              \
            if (r <= g * 256 + 255) {                                                                      \
                T vx[257], vy[257];                                                                        \
                for (int l = 0; l < 256; l++) { vx[l] = M[(size_t)(g * 256 + l) * ld + r]; vy[l] = (T)(g * 256 + l); } \
                const int loopLimit = 256;                                                                 \
                int lim = (r >= g * 256) ? loopLimit - (r % 256) : loopLimit;                              \
                if (lim % 2 != 0) { vx[loopLimit] = 0; vy[loopLimit] = 0; lim++; }                         \
                for (int i = lim >> 1; i > 0; i >>= 1) {                                                   \
                    for (int l = 0; l < i; l++)                                                            \
                        if (ABS(vx[l + i]) > ABS(vx[l]) && g * 256 + l >= r) { vx[l] = vx[l + i]; vy[l] = vy[l + i]; } \
                    if (i % 2 != 0 && i != 1) i++;                                                         \
                }                                                                                          \
                const int sel = (r >= g * 256) ? r % 256 : 0;                                              \
                ox = vx[sel]; oy = vy[sel];                                                                \
            }                                                                                              \
            if (ABS(ox) > ABS(bestv)) { bestv = ox; besti = oy; }                                          \
        }                                                                                                  \
        *value = bestv;                                                                                    \
        return (int)besti;                                                                                 \
    }
DEFINE_QUIRK(float, f32, fabsf)
DEFINE_QUIRK(double, f64, fabs) /* matrix_inversion_FP64.cpp:43-117: the same two kernels on double2 */

/* ---------------------------------------------------------------- synthetic inputs (SURVEY s.8d) */

static inline uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* U[0,100) FP32, stateless: the device generator (csrc/generate.cu) computes the same bits.
 * Distribution per matrix_inv_pyopencl.py:17 / matrix_inv_numpy.py:40. */
static inline float gj_u100(uint64_t seed, uint64_t idx) {
    return (float)(splitmix64(seed ^ idx) >> 40) * (1.0f / 16777216.0f) * 100.0f;
}

/* kind 0: random-uniform; kind 1: diagonally dominant (FP32 row sum in j order). */
void gj_generate_f32(float *A, int n, uint64_t seed, int kind) {
    const uint64_t N = (uint64_t)n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        float *row = A + (size_t)i * N;
        for (int j = 0; j < n; j++) row[j] = gj_u100(seed, (uint64_t)i * N + (uint64_t)j);
        if (kind == 1) {
            float s = 0.0f;
            for (int j = 0; j < n; j++)
                if (j != i) s = s + row[j];
            row[i] = (s + row[i]) + 1.0f;
        }
    }
}

/* Hollow integer fixture of SOL/main_file.cpp:41-52 with MSVC's rand():
 * seed = seed*214013 + 2531011; rand = (seed >> 16) & 0x7fff; initial seed 1. */
void gj_generate_hollow_f32(float *A, int n, uint32_t *state) {
    uint32_t s = *state;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j) { A[(size_t)i * n + j] = 0.0f; continue; }
            s = s * 214013u + 2531011u;
            A[(size_t)i * n + j] = (float)(((s >> 16) & 0x7fff) % 10);
        }
    *state = s;
}

/* ---------------------------------------------------------------- templated bodies */

#define DEFINE_GJ(T, SUF, FMA, FABS, ISFIN)                                                       \
    static inline T elim_##SUF(T a, T c, T u, int nofma) {                                        \
        if (nofma) { volatile T prod = c * u; return a - prod; }                                  \
        return FMA(-c, u, a);                                                                     \
    }                                                                                             \
    /* pivot rule (A.2): scan upward from r, replace only on strict '>' => lowest index on ties; \
     * a NaN candidate never wins, a NaN incumbent is never displaced. */                         \
    static inline int argmax_col_##SUF(const T *M, size_t ld, int n, int r, int col) {            \
        int p = r;                                                                                \
        T best = FABS(M[(size_t)r * ld + col]);                                                   \
        for (int i = r + 1; i < n; i++) {                                                         \
            T v = FABS(M[(size_t)i * ld + col]);                                                  \
            if (v > best) { best = v; p = i; }                                                    \
        }                                                                                         \
        return p;                                                                                 \
    }                                                                                             \
    static int scan_finite_##SUF(const T *X, size_t cnt) {                                        \
        int bad = 0;                                                                              \
        for (size_t k = 0; k < cnt; k++) bad |= !ISFIN(X[k]);                                     \
        return bad ? -1 : 0;                                                                      \
    }                                                                                             \
                                                                                                  \
    /* A.1: the reference's data layout and step order, N x 2N ping-pong collapsed to one        \
     * buffer (the ping-pong only avoids a device race).  forced_piv != NULL replays a given     \
     * pivot sequence instead of searching (used for the FP64 replay of the FP32 pivots). */      \
    int gj_aug_##SUF(const T *A, int n, T *X, int *piv, const int *forced_piv, int flags) {       \
        const int nofma = flags & GJ_NOFMA;                                                       \
        const size_t ld = 2 * (size_t)n;                                                          \
        T *M = (T *)malloc(sizeof(T) * ld * n);                                                   \
        int info = 0;                                                                             \
        for (int i = 0; i < n; i++) {                                                             \
            for (int j = 0; j < n; j++) M[i * ld + j] = A[(size_t)i * n + j];                     \
            for (int j = 0; j < n; j++) M[i * ld + n + j] = (i == j) ? (T)1 : (T)0;               \
        }                                                                                         \
        for (int r = 0; r < n; r++) {                                                             \
            int p = forced_piv ? forced_piv[r] : ((flags & GJ_NOPIVOT) ? r : argmax_col_##SUF(M, ld, n, r, r)); \
            T v = M[(size_t)p * ld + r];                                                          \
            QUIRK_##SUF                                                                           \
            if (piv) piv[r] = p;                                                                  \
            if (v == (T)0 || !ISFIN(v)) { info = r + 1; break; }                                  \
            if (p != r)                                                                           \
                for (size_t j = 0; j < ld; j++) {                                                 \
                    T t = M[r * ld + j]; M[r * ld + j] = M[p * ld + j]; M[p * ld + j] = t;        \
                }                                                                                 \
            T *Mr = M + (size_t)r * ld;                                                           \
            for (size_t j = 0; j < ld; j++) Mr[j] = Mr[j] / v;                                    \
            _Pragma("omp parallel for schedule(static)")                                          \
            for (int i = 0; i < n; i++) {                                                         \
                T *Mi = M + (size_t)i * ld;                                                       \
                T c = Mi[r];                                                                      \
                if (i == r || c == (T)0) continue; /* LIB/mat_inv_32.cpp:28 guard */              \
                for (size_t j = 0; j < ld; j++) Mi[j] = elim_##SUF(Mi[j], c, Mr[j], nofma);       \
            }                                                                                     \
        }                                                                                         \
        if (!info) {                                                                              \
            for (int i = 0; i < n; i++)                                                           \
                for (int j = 0; j < n; j++) X[(size_t)i * n + j] = M[i * ld + n + j];             \
            /* SOL identity check (exact compare) folded into the flag rule + finite scan. */    \
            info = scan_finite_##SUF(X, (size_t)n * n);                                           \
        }                                                                                         \
        free(M);                                                                                  \
        return info;                                                                              \
    }                                                                                             \
                                                                                                  \
    /* A.3: in-place, 2N^3 flops; what the unblocked CUDA path computes step by step.  The      \
     * 'c != 0' guard of fixColumnKernel is dropped here exactly as in the kernels: for finite   \
     * data fma(-0,u,a) == a numerically, only the SIGN of an exact zero can differ from A.1. */  \
    int gj_inplace_##SUF(const T *A, int n, T *X, int *piv_out, int flags) {                      \
        const int nofma = flags & GJ_NOFMA;                                                       \
        const size_t ld = (size_t)n;                                                              \
        int *piv = (int *)malloc(sizeof(int) * (size_t)n);                                        \
        int info = 0;                                                                             \
        if (X != A) memcpy(X, A, sizeof(T) * ld * n);                                             \
        for (int r = 0; r < n; r++) {                                                             \
            int p = (flags & GJ_NOPIVOT) ? r : argmax_col_##SUF(X, ld, n, r, r);                  \
            T v = X[(size_t)p * ld + r];                                                          \
            piv[r] = p;                                                                           \
            if (v == (T)0 || !ISFIN(v)) { info = r + 1; break; }                                  \
            if (p != r)                                                                           \
                for (size_t j = 0; j < ld; j++) {                                                 \
                    T t = X[r * ld + j]; X[r * ld + j] = X[p * ld + j]; X[p * ld + j] = t;        \
                }                                                                                 \
            T *Xr = X + (size_t)r * ld;                                                           \
            const T inv = (T)1 / v;                                                               \
            for (int j = 0; j < n; j++) Xr[j] = Xr[j] / v;                                        \
            Xr[r] = inv;                                                                          \
            _Pragma("omp parallel for schedule(static)")                                          \
            for (int i = 0; i < n; i++) {                                                         \
                if (i == r) continue;                                                             \
                T *Xi = X + (size_t)i * ld;                                                       \
                const T c = Xi[r];                                                                \
                for (int j = 0; j < r; j++) Xi[j] = elim_##SUF(Xi[j], c, Xr[j], nofma);           \
                Xi[r] = elim_##SUF((T)0, c, inv, nofma);                                          \
                for (int j = r + 1; j < n; j++) Xi[j] = elim_##SUF(Xi[j], c, Xr[j], nofma);       \
            }                                                                                     \
        }                                                                                         \
        if (!info) {                                                                              \
            for (int r = n - 1; r >= 0; r--) {                                                    \
                int p = piv[r];                                                                   \
                if (p == r) continue;                                                             \
                for (int i = 0; i < n; i++) {                                                     \
                    T t = X[i * ld + r]; X[i * ld + r] = X[i * ld + p]; X[i * ld + p] = t;        \
                }                                                                                 \
            }                                                                                     \
            info = scan_finite_##SUF(X, ld * n);                                                  \
        }                                                                                         \
        if (piv_out) memcpy(piv_out, piv, sizeof(int) * (size_t)n);                               \
        free(piv);                                                                                \
        return info;                                                                              \
    }                                                                                             \
                                                                                                  \
    /* A.4: right-looking blocked form with an NB-wide panel that is itself factored in          \
     * W-wide sub-panels -- the exact schedule of csrc/gj_blocked.cu: sub-panel factor ->        \
     * (swaps + recurrence + rank-W update) of the rest of the panel -> (swaps + recurrence +    \
     * rank-NB update) of all other columns.  Every element sees the same FMA chain as A.3. */   \
    int gj_blocked_##SUF(const T *A, int n, T *X, int *piv_out, int nb, int w, int flags) {       \
        const int nofma = flags & GJ_NOFMA;                                                       \
        const size_t ld = (size_t)n;                                                              \
        int *piv = (int *)malloc(sizeof(int) * (size_t)n);                                        \
        T *C = (T *)malloc(sizeof(T) * (size_t)nb * n);   /* C[t][i]: multiplier of row i, step t */ \
        T *U = (T *)malloc(sizeof(T) * (size_t)nb * n);   /* U[t][j]: pivot-row snapshot, step t */  \
        T *pv = (T *)malloc(sizeof(T) * (size_t)nb);                                              \
        int info = 0;                                                                             \
        if (w <= 0 || w > nb) w = nb;                                                             \
        if (X != A) memcpy(X, A, sizeof(T) * ld * n);                                             \
        for (int k0 = 0; k0 < n && !info; k0 += nb) {                                             \
            const int kb = (n - k0 < nb) ? n - k0 : nb;                                           \
            /* ---- panel: columns [k0, k0+kb), all n rows, in sub-panels of width w */          \
            for (int s0 = 0; s0 < kb && !info; s0 += w) {                                         \
                const int sw = (kb - s0 < w) ? kb - s0 : w;                                       \
                /* (i) unblocked factor of sub-panel columns [k0+s0, k0+s0+sw) */                \
                for (int t = s0; t < s0 + sw; t++) {                                              \
                    const int r = k0 + t;                                                         \
                    int p = argmax_col_##SUF(X, ld, n, r, r);                                     \
                    T v = X[(size_t)p * ld + r];                                                  \
                    piv[r] = p; pv[t] = v;                                                        \
                    if (v == (T)0 || !ISFIN(v)) { info = r + 1; break; }                          \
                    if (p != r) {                                                                 \
                        for (int j = k0 + s0; j < k0 + s0 + sw; j++) {                            \
                            T x = X[r * ld + j]; X[r * ld + j] = X[p * ld + j]; X[p * ld + j] = x; \
                        }                                                                         \
                        for (int q = s0; q < t; q++) { /* already-recorded multipliers */        \
                            T x = C[(size_t)q * n + r]; C[(size_t)q * n + r] = C[(size_t)q * n + p]; C[(size_t)q * n + p] = x; \
                        }                                                                         \
                    }                                                                             \
                    T *Xr = X + (size_t)r * ld;                                                   \
                    const T inv = (T)1 / v;                                                       \
                    for (int j = k0 + s0; j < k0 + s0 + sw; j++) Xr[j] = Xr[j] / v;               \
                    Xr[r] = inv;                                                                  \
                    for (int i = 0; i < n; i++) {                                                 \
                        T *Xi = X + (size_t)i * ld;                                               \
                        if (i == r) { C[(size_t)t * n + i] = (T)0; continue; }                    \
                        const T c = Xi[r];                                                        \
                        C[(size_t)t * n + i] = c;                                                 \
                        for (int j = k0 + s0; j < k0 + s0 + sw; j++)                              \
                            if (j != r) Xi[j] = elim_##SUF(Xi[j], c, Xr[j], nofma);               \
                        Xi[r] = elim_##SUF((T)0, c, inv, nofma);                                  \
                    }                                                                             \
                }                                                                                 \
                if (info) break;                                                                  \
                /* (ii) rest of the panel (columns of the panel outside the sub-panel) AND the   \
                 * earlier multiplier columns C[0..s0): swaps, recurrence, rank-sw update. */    \
                for (int t = s0; t < s0 + sw; t++) {                                              \
                    const int r = k0 + t, p = piv[r];                                             \
                    if (p == r) continue;                                                         \
                    for (int j = k0; j < k0 + kb; j++) {                                          \
                        if (j >= k0 + s0 && j < k0 + s0 + sw) continue;                           \
                        T x = X[r * ld + j]; X[r * ld + j] = X[p * ld + j]; X[p * ld + j] = x;    \
                    }                                                                             \
                    for (int q = 0; q < s0; q++) {                                                \
                        T x = C[(size_t)q * n + r]; C[(size_t)q * n + r] = C[(size_t)q * n + p]; C[(size_t)q * n + p] = x; \
                    }                                                                             \
                }                                                                                 \
                for (int j = k0; j < k0 + kb; j++) {                                              \
                    if (j >= k0 + s0 && j < k0 + s0 + sw) continue;                               \
                    for (int t = s0; t < s0 + sw; t++) { /* recurrence on the sw pivot rows */   \
                        const int r = k0 + t;                                                     \
                        const T u = X[(size_t)r * ld + j] / pv[t];                                \
                        U[(size_t)t * n + j] = u;                                                 \
                        X[(size_t)r * ld + j] = u;                                                \
                        for (int t2 = s0; t2 < s0 + sw; t2++) {                                   \
                            if (t2 == t) continue;                                                \
                            const int r2 = k0 + t2;                                               \
                            X[(size_t)r2 * ld + j] = elim_##SUF(X[(size_t)r2 * ld + j], C[(size_t)t * n + r2], u, nofma); \
                        }                                                                         \
                    }                                                                             \
                }                                                                                 \
                /* rank-sw trailing update inside the panel, row by row (same per-element chain) */ \
                _Pragma("omp parallel for schedule(static)")                                      \
                for (int i = 0; i < n; i++) {                                                     \
                    if (i >= k0 + s0 && i < k0 + s0 + sw) continue;                               \
                    T *Xi = X + (size_t)i * ld;                                                   \
                    for (int j = k0; j < k0 + kb; j++) {                                          \
                        if (j >= k0 + s0 && j < k0 + s0 + sw) continue;                           \
                        T acc = Xi[j];                                                            \
                        for (int t = s0; t < s0 + sw; t++)                                        \
                            acc = elim_##SUF(acc, C[(size_t)t * n + i], U[(size_t)t * n + j], nofma); \
                        Xi[j] = acc;                                                              \
                    }                                                                             \
                }                                                                                 \
            }                                                                                     \
            if (info) break;                                                                      \
            /* ---- all other columns: swaps, row-block recurrence, trailing update */           \
            for (int t = 0; t < kb; t++) {                                                        \
                const int r = k0 + t, p = piv[r];                                                 \
                if (p == r) continue;                                                             \
                for (int j = 0; j < n; j++) {                                                     \
                    if (j >= k0 && j < k0 + kb) continue;                                         \
                    T x = X[r * ld + j]; X[r * ld + j] = X[p * ld + j]; X[p * ld + j] = x;        \
                }                                                                                 \
            }                                                                                     \
            /* row-block recurrence: per column j the steps t run in order (snapshot, then the other pivot rows);   \
             * columns are independent, so they are processed in chunks with the row accesses contiguous */         \
            _Pragma("omp parallel for schedule(dynamic, 1)")                                      \
            for (int jc = 0; jc < n; jc += 256) {                                                 \
                const int je = (jc + 256 < n) ? jc + 256 : n;                                     \
                for (int t = 0; t < kb; t++) {                                                    \
                    T *Xr = X + (size_t)(k0 + t) * ld;                                            \
                    T *Ut = U + (size_t)t * n;                                                    \
                    const T pvt = pv[t];                                                          \
                    for (int j = jc; j < je; j++) {                                               \
                        if (j >= k0 && j < k0 + kb) continue;                                     \
                        const T u = Xr[j] / pvt;                                                  \
                        Ut[j] = u;                                                                \
                        Xr[j] = u;                                                                \
                    }                                                                             \
                    for (int t2 = 0; t2 < kb; t2++) {                                             \
                        if (t2 == t) continue;                                                    \
                        T *X2 = X + (size_t)(k0 + t2) * ld;                                       \
                        const T c = C[(size_t)t * n + k0 + t2];                                   \
                        for (int j = jc; j < je; j++) {                                           \
                            if (j >= k0 && j < k0 + kb) continue;                                 \
                            X2[j] = elim_##SUF(X2[j], c, Ut[j], nofma);                           \
                        }                                                                         \
                    }                                                                             \
                }                                                                                 \
            }                                                                                     \
            /* per element: acc = X[i][j]; for t: acc = elim(acc, C[t][i], U[t][j]) -- the chain of A.3.  Loop order  \
             * (row, column chunk, step, column) keeps the chunk in L1 and lets the compiler vectorise over j; the    \
             * order of the operations on any ONE element is unchanged, so the result is bit-identical. */            \
            _Pragma("omp parallel for schedule(dynamic, 16)")                                     \
            for (int i = 0; i < n; i++) {                                                         \
                if (i >= k0 && i < k0 + kb) continue;                                             \
                T *Xi = X + (size_t)i * ld;                                                       \
                T ci[256];                                                                        \
                for (int t = 0; t < kb; t++) ci[t] = C[(size_t)t * n + i];                        \
                for (int seg = 0; seg < 2; seg++) {                                               \
                    const int ja = seg ? k0 + kb : 0, jb = seg ? n : k0;                          \
                    for (int jc = ja; jc < jb; jc += 1024) {                                      \
                        const int je = (jc + 1024 < jb) ? jc + 1024 : jb;                         \
                        for (int t = 0; t < kb; t++) {                                            \
                            const T c = ci[t];                                                    \
                            const T *Ut = U + (size_t)t * n;                                      \
                            if (nofma) { for (int j = jc; j < je; j++) Xi[j] = elim_##SUF(Xi[j], c, Ut[j], 1); } \
                            else {                                                                \
                                _Pragma("omp simd")                                               \
                                for (int j = jc; j < je; j++) Xi[j] = FMA(-c, Ut[j], Xi[j]);      \
                            }                                                                     \
                        }                                                                         \
                    }                                                                             \
                }                                                                                 \
            }                                                                                     \
        }                                                                                         \
        if (!info) {                                                                              \
            _Pragma("omp parallel for schedule(static)")                                          \
            for (int i = 0; i < n; i++) {   /* the same transpositions, in the same order, one row at a time */ \
                T *Xi = X + (size_t)i * ld;                                                       \
                for (int r = n - 1; r >= 0; r--) {                                                \
                    const int p = piv[r];                                                         \
                    if (p == r) continue;                                                         \
                    T t = Xi[r]; Xi[r] = Xi[p]; Xi[p] = t;                                        \
                }                                                                                 \
            }                                                                                     \
            info = scan_finite_##SUF(X, ld * n);                                                  \
        }                                                                                         \
        if (piv_out) memcpy(piv_out, piv, sizeof(int) * (size_t)n);                               \
        free(piv); free(C); free(U); free(pv);                                                    \
        return info;                                                                              \
    }

#define QUIRK_f32                                                                         \
    if ((flags & GJ_QUIRK) && !forced_piv) {                                              \
        if (n % 256 != 0) { free(M); return -2; }                                         \
        float qv;                                                                         \
        p = quirk_pivot_f32(M, ld, n, r, &qv);                                            \
        v = qv;                                                                           \
    }
#define QUIRK_f64                                                                         \
    if ((flags & GJ_QUIRK) && !forced_piv) {                                              \
        if (n % 256 != 0) { free(M); return -2; }                                         \
        double qv;                                                                        \
        p = quirk_pivot_f64(M, ld, n, r, &qv);                                            \
        v = qv;                                                                           \
    }
DEFINE_GJ(float, f32, fmaf, fabsf, isfinite)
DEFINE_GJ(double, f64, fma, fabs, isfinite)

/* ---------------------------------------------------------------- metrics (FP64 accumulation) */

/* north_star gate: ||A X - I||_F / (N ||A||_F ||X||_F); also returns the reference's own
 * "Frobenius defect" sqrt(N) - ||A X||_F (SOL/matrix_multiply.cpp:194-200) through *defect. */
double gj_residual_f32(const float *A, const float *X, int n, double *defect) {
    const size_t N = (size_t)n;
    double r2 = 0.0, p2 = 0.0, a2 = 0.0, x2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : r2, p2, a2, x2)
    for (int i = 0; i < n; i++) {
        double *row = (double *)calloc(N, sizeof(double));
        for (size_t k = 0; k < N; k++) {
            const double a = (double)A[i * N + k];
            const float *Xk = X + k * N;
            a2 += a * a;
            for (size_t j = 0; j < N; j++) row[j] += a * (double)Xk[j];
        }
        for (size_t j = 0; j < N; j++) {
            const double x = (double)X[i * N + j];
            x2 += x * x;
            p2 += row[j] * row[j];
            const double d = row[j] - (((size_t)i == j) ? 1.0 : 0.0);
            r2 += d * d;
        }
        free(row);
    }
    if (defect) *defect = sqrt((double)n) - sqrt(p2);
    return sqrt(r2) / ((double)n * sqrt(a2) * sqrt(x2));
}

int gj_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
