// Bring-up / regression check of the 3xTF32 tcgen05 trailing update without Python (starts in seconds on a fresh box).
//   1. pattern tests: one-hot operands whose product is known exactly -> decodes operand-layout mistakes
//   2. random test at npad = 512 against an FP64 host reference and against the FP32 SIMT kernel
//   3. timing of both kernels at npad = 16384
//   4. whole inversions with and without MATINV_FLAG_TF32X3 (gate estimate vs true residual, pivot differences, time)
// Build: g++ -O2 -std=c++17 tools/tc_check.cpp -Iinclude -I/usr/local/cuda/include -Lgpu_matrix_inversion_b200 -lmatinv32
//        -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,'$ORIGIN/../gpu_matrix_inversion_b200' -o tools/tc_check
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "matinv_shim.h"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float urand() {  // U(-1, 1)
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (float)((double)(rng_state >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

#define CUDA_OK(x)                                                                     \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

struct Dev {
    float *W = nullptr, *C = nullptr, *U = nullptr;
    int npad = 0;
    void alloc(int n) {
        npad = n;
        CUDA_OK(cudaMalloc(&W, (size_t)n * n * 4));
        CUDA_OK(cudaMalloc(&C, (size_t)128 * n * 4));
        CUDA_OK(cudaMalloc(&U, (size_t)128 * n * 4));
    }
    void release() { cudaFree(W); cudaFree(C); cudaFree(U); W = C = U = nullptr; }
};

// runs one update in `mode` on a fresh copy of Win, returns the result
static std::vector<float> run_update(Dev &d, const std::vector<float> &Win, const std::vector<float> &C, const std::vector<float> &U,
                                     int k0, int mode) {
    const int n = d.npad;
    CUDA_OK(cudaMemcpy(d.W, Win.data(), (size_t)n * n * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d.C, C.data(), (size_t)128 * n * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d.U, U.data(), (size_t)128 * n * 4, cudaMemcpyHostToDevice));
    const int rc = matinv_debug_trailing_update(d.W, n, n, k0, d.C, d.U, mode, 1, nullptr, nullptr);
    if (rc != 0) {
        printf("matinv_debug_trailing_update(mode %d) -> %d: %s\n", mode, rc, matinv_last_error());
        exit(3);
    }
    std::vector<float> out((size_t)n * n);
    CUDA_OK(cudaMemcpy(out.data(), d.W, (size_t)n * n * 4, cudaMemcpyDeviceToHost));
    return out;
}

static int pattern_test(Dev &d, const char *name, int kk, int amode, int bmode) {
    // A[i][k] = (k == kk) * fa(i),  B[k][j] = (k == kk) * fb(j)  ->  D[i][j] = fa(i) * fb(j); W starts at 0 -> W = -D
    const int n = d.npad, k0 = 128;
    std::vector<float> W((size_t)n * n, 0.0f), C((size_t)128 * n, 0.0f), U((size_t)128 * n, 0.0f);
    auto fa = [&](int i) { return amode ? (float)(i % 251 + 1) : 1.0f; };
    auto fb = [&](int j) { return bmode ? (float)(j % 241 + 1) : 1.0f; };
    for (int i = 0; i < n; i++) C[(size_t)kk * n + i] = fa(i);
    for (int j = 0; j < n; j++) U[(size_t)kk * n + j] = fb(j);
    std::vector<float> out = run_update(d, W, C, U, k0, 1);
    long long bad = 0;
    int fi = -1, fj = -1;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            const bool skipped = (i / 128 == k0 / 128) || (j / 128 == k0 / 128);
            const float want = skipped ? 0.0f : -fa(i) * fb(j);
            if (out[(size_t)i * n + j] != want) {
                if (bad == 0) { fi = i; fj = j; }
                bad++;
            }
        }
    printf("pattern %-28s k=%3d : %s (%lld wrong)\n", name, kk, bad ? "FAIL" : "ok", bad);
    if (bad) {
        printf("  first wrong at (%d,%d): got %g want %g\n", fi, fj, out[(size_t)fi * n + fj], -fa(fi) * fb(fj));
        printf("  -W[0..7][0..7] (want fa(i)*fb(j)):\n");
        for (int i = 0; i < 8; i++) {
            printf("   ");
            for (int j = 0; j < 8; j++) printf(" %8g", -out[(size_t)i * n + j]);
            printf("\n");
        }
        printf("  -W[i][0], i = 0,1,7,8,9,15,16,31,32,64,127: ");
        const int is[] = {0, 1, 7, 8, 9, 15, 16, 31, 32, 64, 127};
        for (int i : is) printf(" %g", -out[(size_t)i * n + 0]);
        printf("\n  -W[0][j], j = 0,1,3,4,7,8,9,15,16,31,32,64,127: ");
        const int js[] = {0, 1, 3, 4, 7, 8, 9, 15, 16, 31, 32, 64, 127};
        for (int j : js) printf(" %g", -out[(size_t)0 * n + j]);
        printf("\n");
    }
    return bad != 0;
}

int main(int argc, char **argv) {
    const int big = (argc > 1) ? atoi(argv[1]) : 16384;
    const bool tc_only = argc > 2 && !strcmp(argv[2], "tc");    // only the tf32x3 inversion at n = big (profiling runs)
    const bool inv_only = tc_only || (argc > 2 && !strcmp(argv[2], "inv"));  // only section 4 at n = big
    if (matinv_device_count() <= 0) { printf("no CUDA device\n"); return 1; }
    int fails = 0;
    Dev d;
    if (!inv_only) {
    d.alloc(512);
    // ---- 1. patterns
    const int kks[] = {0, 1, 3, 4, 7, 8, 15, 16, 17, 127};
    for (int kk : kks) fails += pattern_test(d, "ones x ones", kk, 0, 0);
    fails += pattern_test(d, "row-index x ones", 0, 1, 0);
    fails += pattern_test(d, "ones x col-index", 0, 0, 1);
    fails += pattern_test(d, "row-index x col-index", 21, 1, 1);

    // ---- 2. random, against FP64 host reference
    {
        const int n = 512, k0 = 256;
        std::vector<float> W((size_t)n * n), C((size_t)128 * n), U((size_t)128 * n);
        for (auto &x : W) x = urand() * 50.0f;
        for (auto &x : C) x = urand();
        for (auto &x : U) x = urand() * 100.0f;
        std::vector<float> s = run_update(d, W, C, U, k0, 0), t = run_update(d, W, C, U, k0, 1);
        double e_s = 0, e_t = 0, e_ts = 0, scale = 0;
        long long touched_skip = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                const size_t o = (size_t)i * n + j;
                if (i / 128 == k0 / 128 || j / 128 == k0 / 128) {
                    if (t[o] != W[o]) touched_skip++;
                    continue;
                }
                double acc = 0, mag = fabs((double)W[o]);
                for (int k = 0; k < 128; k++) {
                    const double p = (double)C[(size_t)k * n + i] * (double)U[(size_t)k * n + j];
                    acc += p;
                    mag += fabs(p);
                }
                const double ref = (double)W[o] - acc;
                e_s = fmax(e_s, fabs(s[o] - ref) / mag);
                e_t = fmax(e_t, fabs(t[o] - ref) / mag);
                e_ts = fmax(e_ts, fabs((double)t[o] - (double)s[o]) / mag);
                scale = fmax(scale, mag);
            }
        printf("random n=512: max err / (|w| + sum|c u|):  simt %.3e   tf32x3 %.3e   tf32x3-vs-simt %.3e   (2^-24 = %.3e); skipped tiles touched: %lld\n",
               e_s, e_t, e_ts, ldexp(1.0, -24), touched_skip);
        if (!(e_t < 4e-6) || touched_skip) { printf("random test FAIL\n"); fails++; }
    }
    d.release();
    if (fails) { printf("RESULT: FAIL (%d)\n", fails); return 1; }

    // ---- 3. timing
    {
        const int n = big;
        d.alloc(n);
        std::vector<float> C((size_t)128 * n), U((size_t)128 * n);
        for (auto &x : C) x = urand() * 1e-3f;
        for (auto &x : U) x = urand();
        CUDA_OK(cudaMemset(d.W, 0, (size_t)n * n * 4));
        CUDA_OK(cudaMemcpy(d.C, C.data(), (size_t)128 * n * 4, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(d.U, U.data(), (size_t)128 * n * 4, cudaMemcpyHostToDevice));
        for (int mode = 0; mode < 2; mode++) {
            double ms = 0;
            matinv_debug_trailing_update(d.W, n, n, n / 2, d.C, d.U, mode, 2, &ms, nullptr);  // warm-up
            const int rc = matinv_debug_trailing_update(d.W, n, n, n / 2, d.C, d.U, mode, 20, &ms, nullptr);
            const double m = (double)n - 128;
            printf("timing n=%d mode %d (%s): rc %d  %.3f ms per update  = %.1f TFLOP/s (2 m^2 128), W traffic %.2f TB/s\n", n, mode,
                   mode ? "tf32x3 tcgen05 incl. split" : "fp32 simt", rc, ms, 2.0 * m * m * 128 / (ms * 1e-3) / 1e12,
                   8.0 * m * m / (ms * 1e-3) / 1e12);
        }
        d.release();
    }

    }
    // ---- 4. whole inversions
    const int sizes[] = {1024, 4096, big};
    for (int n : sizes) {
        if (inv_only && n != big) continue;
        float *A = nullptr, *X = nullptr;
        int *piv = nullptr;
        CUDA_OK(cudaMalloc(&A, (size_t)n * n * 4));
        CUDA_OK(cudaMalloc(&X, (size_t)n * n * 4));
        CUDA_OK(cudaMalloc(&piv, (size_t)n * 4));
        matinv_generate_f32_dev(A, n, n, 0xB2000000ull + n, 0, 0, n, nullptr);
        std::vector<int> p0(n), p1(n);
        for (int pass = tc_only ? 1 : 0; pass < 2; pass++) {
            const int flags = pass ? MATINV_FLAG_TF32X3 : 0;
            int rc = matinv_invert_f32_dev(A, n, X, piv, nullptr, flags);  // warm-up (allocations)
            const auto t0 = std::chrono::steady_clock::now();
            rc = matinv_invert_f32_dev(A, n, X, piv, nullptr, flags);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            CUDA_OK(cudaMemcpy((pass ? p1 : p0).data(), piv, (size_t)n * 4, cudaMemcpyDeviceToHost));
            double est = -1;
            int fb = 0;
            matinv_tf32x3_status(&est, &fb, nullptr, nullptr);
            double r[3] = {0, 0, 0};
            double res = -1;
            if (n <= 4096 && rc == 0) {
                matinv_residual_f32_dev(A, X, n, r, nullptr);
                res = sqrt(r[0]) / ((double)n * sqrt(r[1]) * sqrt(r[2]));
            }
            printf("invert n=%5d %-7s: rc %d  %.2f ms wall (%.1f TFLOP/s at 2n^3)  residual %.3e", n, pass ? "tf32x3" : "fp32", rc,
                   sec * 1e3, 2.0 * n * (double)n * n / sec / 1e12, res);
            if (pass) printf("  gate estimate %.3e  fell back %d", est, fb);
            printf("  %s\n", rc ? matinv_last_error() : "");
            if (rc != 0 || (pass && fb)) fails++;
        }
        int diff = 0, first = -1;
        for (int i = 0; i < n; i++)
            if (p0[i] != p1[i]) { if (!diff) first = i; diff++; }
        printf("   pivot rows differing fp32 vs tf32x3: %d of %d (first at step %d)\n", diff, n, first);
        cudaFree(A); cudaFree(X); cudaFree(piv);
    }
    printf("RESULT: %s\n", fails ? "FAIL" : "PASS");
    return fails ? 1 : 0;
}
