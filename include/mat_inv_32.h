#pragma once
// Public C++ surface of the library -- one free function, global namespace, <vector> only, so
// that Matlab's clibgen.generateLibraryDefinition parses it exactly like the reference header
// (/root/reference/Matlab/mat_inv_32.h:1-4; usage README.md:33-51).
//
//   matrix_vector : row-major flattened N x N FP32 matrix (element (i,j) at i*N + j)
//   matrix_order  : N
//   returns       : row-major flattened inverse, or an EMPTY vector when N <= 0, the input is
//                   not square, the matrix is singular / non-finite, or no CUDA device exists.
#include <vector>

std::vector<float> matrix_inv_32(std::vector<float> matrix_vector, int matrix_order);
