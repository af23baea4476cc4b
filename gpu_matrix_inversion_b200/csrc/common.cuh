// Shared device helpers for the Gauss-Jordan kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MATINV_NB 128    // panel width == GEMM tile edge
#define MATINV_SUBW 16   // sub-panel width inside a panel
#define MATINV_RB 64     // rows per CTA in the panel-step / argmax kernels

typedef unsigned long long u64;

// Pivot key: max over keys == arg max |x| with the LOWEST row index on ties (north_star rule;
// intent of maxPivotKernel/finalMaxPivotKernel, /root/reference/Matlab/mat_inv_32/mat_inv_32/
// mat_inv_32.cpp:61-132: fabs compare, strict '>').  Layout: [63:32] |x| bits, [31:1]
// 0x7FFFFFFF - row, [0] sign of x -- so the pivot VALUE travels with the key and nobody has to
// re-read the matrix.  NaN: a NaN candidate never wins (fabs(NaN) > x is false, :91,:124) -> its
// magnitude maps to 0; a NaN incumbent (row == first candidate row) is never displaced -> its
// magnitude maps to 0xFFFFFFFF and the value decodes as NaN (=> singular).
__device__ __forceinline__ u64 gj_key(float x, int row, bool incumbent) {
    const float a = fabsf(x);
    const unsigned int b = (a == a) ? __float_as_uint(a) : (incumbent ? 0xFFFFFFFFu : 0u);
    const unsigned int lo = ((0x7FFFFFFFu - (unsigned int)row) << 1) | (__float_as_uint(x) >> 31);
    return ((u64)b << 32) | (u64)lo;
}
// The same key built from its parts (magnitude bits with the NaN rules, row, value for the sign).
__device__ __forceinline__ unsigned gj_mag(float x, bool incumbent) {
    const float a = fabsf(x);
    return (a == a) ? __float_as_uint(a) : (incumbent ? 0xFFFFFFFFu : 0u);
}
__device__ __forceinline__ u64 gj_key_from(unsigned mag, int row, float x) {
    const unsigned int lo = ((0x7FFFFFFFu - (unsigned int)row) << 1) | (__float_as_uint(x) >> 31);
    return ((u64)mag << 32) | (u64)lo;
}
__device__ __forceinline__ int gj_key_row(u64 k) { return (int)(0x7FFFFFFFu - ((unsigned int)(k & 0xFFFFFFFFull) >> 1)); }
__device__ __forceinline__ float gj_key_value(u64 k) {
    const unsigned int b = (unsigned int)(k >> 32);
    if (b == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
    return __uint_as_float(b | ((unsigned int)(k & 1ull) << 31));
}

__device__ __forceinline__ u64 warp_max_u64(u64 k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    return k;
}

__device__ __forceinline__ float f4get(const float4 &v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ void f4set(float4 &v, int k, float x) {
    if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x;
}

// The one elimination primitive: a <- a - c*u as a single FMA (oracle: fmaf(-c, u, a)).
__device__ __forceinline__ float gj_elim(float a, float c, float u) { return fmaf(-c, u, a); }

__device__ __forceinline__ bool gj_bad_pivot(float v) { return v == 0.0f || !isfinite(v); }

// splitmix64 counter generator -- bit-identical to oracle/gj_oracle.c:gj_u100.
__host__ __device__ __forceinline__ float gj_u100(u64 seed, u64 idx) {
    u64 z = (seed ^ idx) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
#ifdef __CUDA_ARCH__
    // explicit rn multiply: keeps nvcc from contracting "* 100.0f" with a following add into an FMA
    return __fmul_rn(__fmul_rn((float)(z >> 40), 1.0f / 16777216.0f), 100.0f);
#else
    return (float)(z >> 40) * (1.0f / 16777216.0f) * 100.0f;
#endif
}
