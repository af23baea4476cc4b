// C++ boundary of the library: the reference's one public function, same signature, same
// argument checks, same error convention -- implemented on the C-ABI shim instead of OpenCL.
//
// Mirrors /root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:
//   :207-209  matrix_order <= 0                      -> {}
//   :212-215  int(size / matrix_order) != matrix_order -> {}   (integer division: size = N*N + k with
//             0 <= k < N is accepted and the tail ignored -- replicated on purpose)
//   :391-394  any device error                        -> {}   (never throws)
// plus the singular-matrix rule that only the development copy implements
// (matrix_inv_solution/.../matrix_inversion_FP32.cpp:814-835: left half of [A|I] != I -> {}),
// which README.md:54 documents as the library contract ("In case of invalid matrix an empty
// vector is returned").
#include "../../include/mat_inv_32.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

#include "../../include/matinv_shim.h"

std::vector<float> matrix_inv_32(std::vector<float> matrix_vector, int matrix_order) {
    if (matrix_order <= 0) return {};
    const int matrix_height = int(matrix_vector.size() / (size_t)matrix_order);
    if (matrix_height != matrix_order) return {};

    const size_t count = (size_t)matrix_order * (size_t)matrix_order;
    // The argument arrives BY VALUE (LIB/mat_inv_32.h:4): this copy is ours, so the inverse is written over it and the vector
    // is moved out -- no second N x N allocation (value-initialising a fresh 1 GiB vector at N = 16384 costs more than
    // the PCIe transfer).  The reference allocates a separate result (LIB:379); the caller cannot tell the difference.
    float *const io = matrix_vector.data();
    int flags = 0;
    const char *verbose = std::getenv("MATINV_VERBOSE");
    if (verbose && verbose[0] && verbose[0] != '0') flags |= MATINV_FLAG_VERBOSE;
    // opt-in, off by default: trailing updates on the tensor cores (3xTF32, residual-gated, not bit-identical to the
    // reference's FMA chain) -- the signature clibgen binds cannot carry a flag, so the switch is an environment variable
    const char *tc = std::getenv("MATINV_TF32X3");
    if (tc && tc[0] && tc[0] != '0') flags |= MATINV_FLAG_TF32X3;
    // opt-in: MATINV_NGPU = k > 1 column-shards one inversion over k GPUs (bit-identical result).  From host memory the
    // transfers dominate below N ~ 32768 (measured with std::vector buffers: N=32768 1.54 s on one GPU, 1.4-1.9 s on 2-8; the
    // device-resident schedule itself scales 7.9 x on 8 GPUs at N=65536), so smaller orders stay on the single-device path
    // unless MATINV_NGPU_MIN_ORDER says otherwise
    const char *ng = std::getenv("MATINV_NGPU");
    const int ngpu = ng ? std::atoi(ng) : 1;
    const char *mo = std::getenv("MATINV_NGPU_MIN_ORDER");
    const int min_order = mo ? std::atoi(mo) : 32768;
    const int rc = (ngpu > 1 && matrix_order >= min_order && !(flags & MATINV_FLAG_TF32X3))
                       ? matinv_invert_sharded_f32(io, matrix_order, io, nullptr, ngpu, 0, flags & MATINV_FLAG_VERBOSE)
                       : matinv_invert_f32(io, matrix_order, io, nullptr, flags);
    if (rc == MATINV_OK) {
        matrix_vector.resize(count);   // drops the ignored tail of an N*N + k input (LIB:212-215)
        return matrix_vector;
    }
    if (rc < 0) std::cerr << "ERRORE N\xC2\xB0: " << rc << " (" << matinv_last_error() << ")" << std::endl;  // LIB:392
    return {};
}
