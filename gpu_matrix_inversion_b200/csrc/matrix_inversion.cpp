// C++ surface of the development copy's function set (include/matrix_inversion.h) on the C-ABI shim.
//
// Argument checks follow the reference:
//   matrix_inversion_FP64.cpp:209-217          order <= 0 -> {};  int(size / order) != order -> {}
//   matrix_inversion_no_pivots.cpp:115-123     the same two checks
//   matrix_inversion_FP32.cpp:814-835, README.md:54   singular / invalid -> {}
// Device errors never propagate as exceptions: they print the reference's "ERRORE N°" line and return {} (or NaN for
// matrix_multiply, which has no empty value to return).
#include "../../include/matrix_inversion.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <utility>

#include "../../include/mat_inv_32.h"
#include "../../include/matinv_shim.h"

namespace {

int env_flags() {
    const char *verbose = std::getenv("MATINV_VERBOSE");
    return (verbose && verbose[0] && verbose[0] != '0') ? MATINV_FLAG_VERBOSE : 0;
}

std::vector<double> invert_f64(std::vector<double> &matrix_vector, int matrix_order, int flags) {
    if (matrix_order <= 0) return {};
    const int matrix_height = int(matrix_vector.size() / (size_t)matrix_order);
    if (matrix_height != matrix_order) return {};
    // the caller's by-value copy is ours: the inverse is written over it and moved out (see mat_inv_32.cpp)
    const int rc = matinv_invert_f64(matrix_vector.data(), matrix_order, matrix_vector.data(), nullptr, flags | env_flags());
    if (rc == MATINV_OK) {
        matrix_vector.resize((size_t)matrix_order * (size_t)matrix_order);
        return std::move(matrix_vector);
    }
    if (rc < 0) std::cerr << "ERRORE N\xC2\xB0: " << rc << " (" << matinv_last_error() << ")" << std::endl;
    return {};
}

}  // namespace

std::vector<float> matrix_inversion_FP32(std::vector<float> matrix_vector, int matrix_order) {
    return matrix_inv_32(std::move(matrix_vector), matrix_order);
}

std::vector<double> matrix_inversion_FP64(std::vector<double> matrix_vector, int matrix_order) {
    return invert_f64(matrix_vector, matrix_order, 0);
}

std::vector<double> matrix_inversion_no_pivots(std::vector<double> matrix_vector, int matrix_order) {
    return invert_f64(matrix_vector, matrix_order, MATINV_FLAG_NOPIVOT);
}


double matrix_multiply(std::vector<double> matriceB, std::vector<double> matriceA) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const int n = (int)std::sqrt((double)matriceA.size());
    if (n <= 0 || (size_t)n * n != matriceA.size() || matriceB.size() != matriceA.size()) return nan;
    double out[4] = {0, 0, 0, 0};
    const int rc = matinv_host_defect_f64(matriceA.data(), matriceB.data(), n, out);
    if (rc != MATINV_OK) {
        std::cerr << "ERRORE N\xC2\xB0: " << rc << " (" << matinv_last_error() << ")" << std::endl;
        return nan;
    }
    const double errore = std::sqrt((double)n) - std::sqrt(out[3]);
    if (env_flags() & MATINV_FLAG_VERBOSE) std::cout << "\nERRORE: " << errore << std::endl;   // matrix_multiply.cpp:202
    return errore;
}
