#pragma once
// The function set of the reference's development copy
// (/root/reference/matrix_inv_solution/matrix_inversion_solution/matrix_inversion/headers.h:5-11), same names, same
// by-value std::vector signatures, same error convention (invalid or singular input -> empty vector), implemented on
// libmatinv32.so (include/matinv_shim.h).  The *_bench variants and res_struct.h are timing wrappers and are not mirrored.
#include <vector>

// sqrt(order) - ||A B||_F for two row-major square matrices of the same order: 0 when B is the inverse of A
// (matrix_multiply.cpp:15-212; the first argument is the right-hand factor, as in the reference).
double matrix_multiply(std::vector<double> matriceB, std::vector<double> matriceA);

// FP32 Gauss-Jordan with partial pivoting (matrix_inversion_FP32.cpp:12); identical to matrix_inv_32.
std::vector<float> matrix_inversion_FP32(std::vector<float> matrix_vector, int matrix_order);

// FP64 Gauss-Jordan with partial pivoting (matrix_inversion_FP64.cpp:13).
std::vector<double> matrix_inversion_FP64(std::vector<double> matrix_vector, int matrix_order);

// FP64 Gauss-Jordan without row interchanges, for matrices whose diagonal never vanishes (matrix_inversion_no_pivots.cpp:10).
std::vector<double> matrix_inversion_no_pivots(std::vector<double> matrix_vector, int matrix_order);
