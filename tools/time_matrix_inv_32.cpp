// Wall-clock of the reference-facing call itself: std::vector<float> matrix_inv_32(std::vector<float>, int) with ordinary
// (pageable) vectors, as main_file.cpp / the clibgen interface use it.   usage: time_matrix_inv_32 [n] [reps]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mat_inv_32.h"

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    std::vector<float> orig((size_t)n * n);
    uint64_t s = 12345;
    for (auto &x : orig) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (float)((s >> 40) % 100000) / 1000.0f; }
    uint64_t h0 = 0;
    for (int r = 0; r < reps + 1; r++) {
        std::vector<float> a = orig;                       // the caller's copy (outside the timed call)
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<float> x = matrix_inv_32(std::move(a), n);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (x.size() != orig.size()) { printf("EMPTY result\n"); return 1; }
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < x.size(); i += 97) { uint32_t b; memcpy(&b, &x[i], 4); h = (h ^ b) * 1099511628211ull; }
        if (r == 0) h0 = h;
        printf("matrix_inv_32 n=%d call %d: %.1f ms (%.1f TFLOP/s at 2N^3)%s hash %016llx%s\n", n, r, dt * 1e3, 2.0 * n * (double)n * n / dt / 1e12,
               r == 0 ? " [first call: context + workspace]" : "", (unsigned long long)h, h == h0 ? "" : " MISMATCH");
    }
    return 0;
}
