// C bridge to the UNMODIFIED reference functions (compiled from /root/reference by oracle/Makefile.ref).
// TEST INFRASTRUCTURE ONLY.  The two prototypes are the reference's own
// (/root/reference/Matlab/mat_inv_32.h:4 and .../matrix_inversion/headers.h:7).
#include <cstring>
#include <vector>

std::vector<float> matrix_inv_32(std::vector<float> matrix_vector, int matrix_order);
std::vector<float> matrix_inversion_FP32(std::vector<float> matrix_vector, int matrix_order);

static int run(std::vector<float> (*f)(std::vector<float>, int), const float *A, long long count, int n, float *X) {
    std::vector<float> in(A, A + count);
    std::vector<float> out = f(in, n);
    if (out.empty()) return 1;
    std::memcpy(X, out.data(), sizeof(float) * (size_t)n * (size_t)n);
    return 0;
}

extern "C" {
// shipped library (no singular check)
int ref_matrix_inv_32(const float *A, long long count, int n, float *X) { return run(matrix_inv_32, A, count, n, X); }
// development copy (identity check => {} on singular input)
int ref_matrix_inversion_FP32(const float *A, long long count, int n, float *X) { return run(matrix_inversion_FP32, A, count, n, X); }
}
