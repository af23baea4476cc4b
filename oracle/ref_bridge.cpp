// C bridge to the UNMODIFIED reference functions (compiled from /root/reference by oracle/Makefile.ref).
// TEST INFRASTRUCTURE ONLY.  The two prototypes are the reference's own
// (/root/reference/Matlab/mat_inv_32.h:4 and .../matrix_inversion/headers.h:7).
#include <cstring>
#include <vector>

std::vector<float> matrix_inv_32(std::vector<float> matrix_vector, int matrix_order);
std::vector<float> matrix_inversion_FP32(std::vector<float> matrix_vector, int matrix_order);
// headers.h:9, :11, :5
std::vector<double> matrix_inversion_FP64(std::vector<double> matrix_vector, int matrix_order);
std::vector<double> matrix_inversion_no_pivots(std::vector<double> matrix_vector, int matrix_order);
double matrix_multiply(std::vector<double> matriceA, std::vector<double> matriceB);

static int run(std::vector<float> (*f)(std::vector<float>, int), const float *A, long long count, int n, float *X) {
    std::vector<float> in(A, A + count);
    std::vector<float> out = f(in, n);
    if (out.empty()) return 1;
    std::memcpy(X, out.data(), sizeof(float) * (size_t)n * (size_t)n);
    return 0;
}

static int run64(std::vector<double> (*f)(std::vector<double>, int), const double *A, long long count, int n, double *X) {
    std::vector<double> in(A, A + count);
    std::vector<double> out = f(in, n);
    if (out.empty()) return 1;
    std::memcpy(X, out.data(), sizeof(double) * (size_t)n * (size_t)n);
    return 0;
}

extern "C" {
int ref_matrix_inversion_FP64(const double *A, long long count, int n, double *X) { return run64(matrix_inversion_FP64, A, count, n, X); }
int ref_matrix_inversion_no_pivots(const double *A, long long count, int n, double *X) { return run64(matrix_inversion_no_pivots, A, count, n, X); }
double ref_matrix_multiply(const double *first, const double *second, int n) {
    std::vector<double> a(first, first + (size_t)n * n), b(second, second + (size_t)n * n);
    return matrix_multiply(a, b);
}
// shipped library (no singular check)
int ref_matrix_inv_32(const float *A, long long count, int n, float *X) { return run(matrix_inv_32, A, count, n, X); }
// development copy (identity check => {} on singular input)
int ref_matrix_inversion_FP32(const float *A, long long count, int n, float *X) { return run(matrix_inversion_FP32, A, count, n, X); }
}
