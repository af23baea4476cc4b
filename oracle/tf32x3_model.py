"""TEST INFRASTRUCTURE -- numpy restatement of the arithmetic of the optional 3xTF32 trailing update
(gpu_matrix_inversion_b200/csrc/gj_gemm_tc.cu).  Only tests/ may import this module.

The reference's trailing update is fixColumnKernel applied column by column
(/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:13-57): w <- w - c*u in FP32.  The tensor-core variant
replaces the 128-term FMA chain by

    hi(x) = rna_tf32(x)            (round to nearest, ties away from zero, to 10 explicit mantissa bits: cvt.rna.tf32.f32)
    lo(x) = rna_tf32(x - hi(x))    (the subtraction is exact in FP32)
    D     = sum_t  lo(c_t) hi(u_t) + hi(c_t) lo(u_t) + hi(c_t) hi(u_t)     (tensor cores, FP32 accumulator)
    w    <- w - D

There is no bit-exact oracle for this path -- the accumulation order inside the tensor core is not specified -- which is
why the product gates its result by residual.  What this model pins is everything that IS specified: the split, which
products are formed, and the size of the terms that are dropped (lo*lo, and the rounding of lo), i.e. the claim that the
variant is FP32-grade: |D_model - c.u| <= 2^-21 * sum|c_t u_t| with the products summed exactly (FP64 here).
"""
from __future__ import annotations

import numpy as np


def rna_tf32(x: np.ndarray) -> np.ndarray:
    """cvt.rna.tf32.f32: keep 10 explicit mantissa bits, round half away from zero (add half an ulp to the magnitude,
    truncate).  Inf/NaN pass through."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32)
    finite = (b & np.uint32(0x7F800000)) != np.uint32(0x7F800000)
    r = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return np.where(finite, r, b).astype(np.uint32).view(np.float32)


def split(x: np.ndarray):
    """(hi, lo) with x ~= hi + lo, both representable in TF32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = rna_tf32(x)
    with np.errstate(invalid="ignore"):
        lo = rna_tf32((x - hi).astype(np.float32))
    return hi, lo


def product_model(C: np.ndarray, U: np.ndarray) -> np.ndarray:
    """D[i][j] = sum_t (lo(C[t][i]) hi(U[t][j]) + hi(C[t][i]) lo(U[t][j]) + hi(C[t][i]) hi(U[t][j])) with the products
    (exact in FP64: 11 x 11 bits) summed in FP64.  C: (K, M), U: (K, N), both K-major like CmT / U on the device."""
    ch, cl = split(C)
    uh, ul = split(U)
    ch, cl, uh, ul = (a.astype(np.float64) for a in (ch, cl, uh, ul))
    return cl.T @ uh + ch.T @ ul + ch.T @ uh


def trailing_update_model(W: np.ndarray, C: np.ndarray, U: np.ndarray) -> np.ndarray:
    """W - D, rounded to FP32 once (the epilogue's single FP32 subtraction of an FP32 accumulator, modelled without the
    accumulator's own rounding)."""
    return (W.astype(np.float64) - product_model(C, U)).astype(np.float32)
