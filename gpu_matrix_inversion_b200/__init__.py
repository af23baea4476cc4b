"""gpu_matrix_inversion_b200 -- Python host of the B200 Gauss-Jordan inverter.

Thin ctypes binding of the C-ABI in include/matinv_shim.h (libmatinv32.so, built in-tree by
`make` / __graft_entry__.build()).  The names mirror the reference's two surfaces:

* ``matrix_inv_32(b, N)``   -- the C++ library function (/root/reference/Matlab/mat_inv_32.h:4):
  row-major flattened input, flattened inverse out, EMPTY result for invalid / singular input.
* ``driver.matrix_inv(file, N)`` -- the PyOpenCL driver (/root/reference/matrix_inv_pyopencl.py:15-352).

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), every
compute call raises / returns the NODEVICE status when no CUDA device is present, and a missing
extension raises ImportError at import time.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

__all__ = ["lib", "matrix_inv_32", "invert", "invert_dev", "invert_batched", "invert_batched_dev", "device_count",
           "last_error", "last_timing", "MatinvError", "OK", "SINGULAR", "FLAG_UNBLOCKED", "FLAG_VERBOSE",
           "FLAG_NOCHECK", "FLAG_TF32X3", "TF32X3_GATE", "tf32x3_status", "debug_trailing_update", "probe_residual_dev", "tf32x3_gate_dev", "TF32X3_GATE_SCALED", "invert_sharded", "sharded_synthetic",
           "invert_batched_ngpu", "EXPORTS"]

OK, SINGULAR = 0, 1
E_INVALID, E_NODEVICE, E_CUDA, E_UNSUPPORTED = -1, -2, -3, -4
FLAG_TF32X3, FLAG_UNBLOCKED, FLAG_VERBOSE, FLAG_NOCHECK, FLAG_NOPIVOT = 1, 2, 4, 8, 16
TF32X3_GATE = 1e-5   # MATINV_TF32X3_GATE (include/matinv_shim.h)
TF32X3_GATE_SCALED = 1e-7   # MATINV_TF32X3_GATE_SCALED: bound on estimate * sqrt(n)

_SO = Path(__file__).resolve().parent / "libmatinv32.so"

# every symbol include/matinv_shim.h declares (tests check the .so exports each one)
EXPORTS = [
    "matinv_device_count", "matinv_last_error", "matinv_shutdown", "matinv_invert_f32", "matinv_invert_f32_dev",
    "matinv_invert_batched_f32", "matinv_invert_batched_f32_dev", "matinv_shard_panel_bytes", "matinv_shard_create",
    "matinv_shard_destroy", "matinv_shard_local", "matinv_shard_set_block", "matinv_shard_get_block",
    "matinv_shard_generate", "matinv_shard_factor",
    "matinv_shard_apply", "matinv_shard_apply_ex", "matinv_shard_status", "matinv_generate_f32_dev", "matinv_generate_batched_f32_dev",
    "matinv_residual_f32_dev", "matinv_last_timing", "matinv_ffma_peak_tflops", "matinv_profile_enable",
    "matinv_profile_read", "matinv_debug_trace", "matinv_invert_f64", "matinv_invert_f64_dev", "matinv_residual_f64_dev",
    "matinv_host_defect_f64", "matinv_tf32x3_status", "matinv_debug_trailing_update", "matinv_probe_residual_f32_dev", "matinv_tf32x3_gate_dev", "matinv_invert_sharded_f32", "matinv_sharded_synthetic_f32",
    "matinv_invert_batched_f32_ngpu", "matinv_nccl_version", "matinv_last_phases", "matinv_debug_pipeline_plan",
]


class MatinvError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"matinv error {code}: {text}")
        self.code = code


def _load() -> ctypes.CDLL:
    if not _SO.exists():
        raise ImportError(
            f"{_SO} is missing: build the CUDA extension first (`make` or `python -c 'import __graft_entry__ as g; "
            f"g.build()'`).  There is no CPU fallback.")
    L = ctypes.CDLL(str(_SO))
    fp, ip, vp, dp = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
    i, ll, ull = ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong
    L.matinv_device_count.restype = i
    L.matinv_last_error.restype = ctypes.c_char_p
    L.matinv_shutdown.restype = None
    L.matinv_invert_f32.argtypes = [fp, i, fp, ip, i]
    L.matinv_invert_f32_dev.argtypes = [fp, i, fp, ip, vp, i]
    L.matinv_invert_batched_f32.argtypes = [fp, i, ll, fp, ip, i]
    L.matinv_invert_batched_f32_dev.argtypes = [fp, i, ll, fp, ip, vp, i]
    L.matinv_generate_f32_dev.argtypes = [fp, i, ll, ull, i, i, i, vp]
    L.matinv_generate_batched_f32_dev.argtypes = [fp, i, ll, ll, ull, vp]
    L.matinv_residual_f32_dev.argtypes = [fp, fp, i, dp, vp]
    L.matinv_invert_f64.argtypes = [fp, i, fp, ip, i]
    L.matinv_invert_f64_dev.argtypes = [fp, i, fp, ip, vp, i]
    L.matinv_residual_f64_dev.argtypes = [fp, fp, i, dp, vp]
    L.matinv_host_defect_f64.argtypes = [fp, fp, i, dp]
    L.matinv_last_timing.argtypes = [dp, dp]
    L.matinv_last_phases.argtypes = [dp]
    L.matinv_last_phases.restype = i
    L.matinv_debug_pipeline_plan.argtypes = [i, ip, ip, ip, ip, ip]
    L.matinv_debug_pipeline_plan.restype = i
    L.matinv_ffma_peak_tflops.argtypes = [dp, vp]
    L.matinv_profile_enable.argtypes = [i]
    L.matinv_profile_enable.restype = None
    L.matinv_profile_read.argtypes = [dp, ctypes.POINTER(ll), dp, ctypes.POINTER(ll)]
    L.matinv_profile_read.restype = i
    L.matinv_debug_trace.argtypes = [i, ctypes.c_void_p]
    L.matinv_debug_trace.restype = i
    L.matinv_shard_panel_bytes.argtypes = [i]
    L.matinv_shard_panel_bytes.restype = ll
    L.matinv_shard_create.argtypes = [i, i, i, ctypes.POINTER(vp)]
    L.matinv_shard_destroy.argtypes = [vp]
    L.matinv_shard_destroy.restype = None
    L.matinv_shard_local.argtypes = [vp, ctypes.POINTER(ll), ctypes.POINTER(ll)]
    L.matinv_shard_local.restype = vp
    L.matinv_shard_set_block.argtypes = [vp, i, vp, ll, vp]
    L.matinv_shard_get_block.argtypes = [vp, i, vp, ll, vp]
    L.matinv_shard_set_block.restype = i
    L.matinv_shard_get_block.restype = i
    L.matinv_shard_generate.argtypes = [vp, ull, i, vp]
    L.matinv_shard_factor.argtypes = [vp, i, vp, vp]
    L.matinv_shard_apply.argtypes = [vp, i, vp, vp]
    L.matinv_shard_apply_ex.argtypes = [vp, i, vp, vp, i, i]
    L.matinv_shard_apply_ex.restype = i
    L.matinv_shard_status.argtypes = [vp, ip, ip, vp]
    L.matinv_tf32x3_status.argtypes = [dp, ip, ctypes.POINTER(ll), ctypes.POINTER(ll)]
    L.matinv_debug_trailing_update.argtypes = [fp, ll, i, i, fp, fp, i, i, dp, vp]
    L.matinv_probe_residual_f32_dev.argtypes = [fp, fp, i, dp, vp]
    L.matinv_tf32x3_gate_dev.argtypes = [fp, fp, i, dp, vp]
    L.matinv_invert_sharded_f32.argtypes = [fp, i, fp, ip, i, i, i]
    L.matinv_sharded_synthetic_f32.argtypes = [i, ull, i, i, ip, dp]
    L.matinv_invert_batched_f32_ngpu.argtypes = [fp, i, ll, fp, ip, i, i]
    L.matinv_nccl_version.argtypes = []
    for name in ("matinv_invert_sharded_f32", "matinv_sharded_synthetic_f32", "matinv_invert_batched_f32_ngpu", "matinv_nccl_version"):
        getattr(L, name).restype = i
    for name in ("matinv_tf32x3_status", "matinv_debug_trailing_update", "matinv_probe_residual_f32_dev", "matinv_tf32x3_gate_dev"):
        getattr(L, name).restype = i
    for name in ("matinv_invert_f32", "matinv_invert_f32_dev", "matinv_invert_batched_f32",
                 "matinv_invert_batched_f32_dev", "matinv_generate_f32_dev", "matinv_generate_batched_f32_dev",
                 "matinv_residual_f32_dev", "matinv_last_timing", "matinv_ffma_peak_tflops", "matinv_shard_create",
                 "matinv_shard_generate", "matinv_shard_factor", "matinv_shard_apply", "matinv_shard_status",
                 "matinv_invert_f64", "matinv_invert_f64_dev", "matinv_residual_f64_dev", "matinv_host_defect_f64"):
        getattr(L, name).restype = i
    return L


lib = _load()


def device_count() -> int:
    return int(lib.matinv_device_count())


def last_error() -> str:
    return lib.matinv_last_error().decode("utf-8", "replace")


def last_timing():
    """(total_s, compute_s) of the last host-pointer inversion on this thread -- the reference's
    'Tempo Totale Impiegato' / 'Tempo Computazione' (mat_inv_32.cpp:385-386)."""
    t, c = ctypes.c_double(), ctypes.c_double()
    if lib.matinv_last_timing(ctypes.byref(t), ctypes.byref(c)) != 0:
        return None
    return t.value, c.value


def last_phases():
    """Phase split of the last host-pointer inversion on this thread, like the reference's Res.times
    (FP32_bench.cpp:256-443): dict(setup, h2d, factor, extract_d2h, total) in seconds, or None."""
    out = (ctypes.c_double * 5)()
    if lib.matinv_last_phases(out) != 0:
        return None
    return dict(zip(("setup", "h2d", "factor", "extract_d2h", "total"), (float(x) for x in out)))


def _check(rc: int) -> int:
    if rc < 0:
        raise MatinvError(rc, last_error())
    return rc


def invert(A: np.ndarray, flags: int = 0, want_piv: bool = False):
    """Invert one n x n FP32 matrix held in host memory.  Returns X (or None when singular) and,
    if want_piv, the pivot vector."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise ValueError("square matrix expected")
    X = np.empty_like(A)
    piv = np.empty(n, dtype=np.int32) if want_piv else None
    rc = _check(lib.matinv_invert_f32(A.ctypes.data, n, X.ctypes.data, piv.ctypes.data if want_piv else None, flags))
    X = X if rc == OK else None
    return (X, piv) if want_piv else X


def matrix_inv_32(matrix_vector, matrix_order: int) -> np.ndarray:
    """Python twin of `std::vector<float> matrix_inv_32(std::vector<float>, int)`: same checks
    (mat_inv_32.cpp:207-215, integer-division squareness test included), same error convention
    (empty result, never raises for bad input / singular matrices / device errors)."""
    v = np.ascontiguousarray(matrix_vector, dtype=np.float32).ravel()
    n = int(matrix_order)
    if n <= 0 or v.size // n != n:
        return np.empty(0, dtype=np.float32)
    flags = FLAG_TF32X3 if os.environ.get("MATINV_TF32X3", "0") not in ("", "0") else 0   # same opt-in as mat_inv_32.cpp
    try:
        X = invert(v[: n * n].reshape(n, n), flags=flags)
    except MatinvError:
        return np.empty(0, dtype=np.float32)
    return np.empty(0, dtype=np.float32) if X is None else X.ravel()


def invert_f64(A: np.ndarray, nopivot: bool = False, want_piv: bool = False, flags: int = 0):
    """FP64 inversion from host memory (device side of matrix_inversion_FP64 / matrix_inversion_no_pivots).  Returns X
    (None when singular) and, if want_piv, the pivot vector (the identity permutation with nopivot)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise ValueError("square matrix expected")
    X = np.empty_like(A)
    piv = np.empty(n, dtype=np.int32) if want_piv else None
    f = flags | (FLAG_NOPIVOT if nopivot else 0)
    rc = _check(lib.matinv_invert_f64(A.ctypes.data, n, X.ctypes.data, piv.ctypes.data if want_piv else None, f))
    X = X if rc == OK else None
    return (X, piv) if want_piv else X


def _vector_f64(matrix_vector, matrix_order: int, nopivot: bool) -> np.ndarray:
    v = np.ascontiguousarray(matrix_vector, dtype=np.float64).ravel()
    n = int(matrix_order)
    if n <= 0 or v.size // n != n:
        return np.empty(0, dtype=np.float64)
    try:
        X = invert_f64(v[: n * n].reshape(n, n), nopivot=nopivot)
    except MatinvError:
        return np.empty(0, dtype=np.float64)
    return np.empty(0, dtype=np.float64) if X is None else X.ravel()


def matrix_inversion_FP64(matrix_vector, matrix_order: int) -> np.ndarray:
    """Python twin of `std::vector<double> matrix_inversion_FP64(std::vector<double>, int)`
    (matrix_inversion_FP64.cpp:13, checks :209-217): empty result for invalid or singular input."""
    return _vector_f64(matrix_vector, matrix_order, False)


def matrix_inversion_no_pivots(matrix_vector, matrix_order: int) -> np.ndarray:
    """Python twin of `matrix_inversion_no_pivots` (matrix_inversion_no_pivots.cpp:10): no row interchanges."""
    return _vector_f64(matrix_vector, matrix_order, True)


def matrix_multiply(matriceB, matriceA) -> float:
    """Python twin of `double matrix_multiply(std::vector<double> B, std::vector<double> A)` (matrix_multiply.cpp:15-212):
    sqrt(order) - ||A B||_F, computed on the device."""
    A = np.ascontiguousarray(matriceA, dtype=np.float64).ravel()
    B = np.ascontiguousarray(matriceB, dtype=np.float64).ravel()
    n = int(round(np.sqrt(A.size)))
    if n <= 0 or n * n != A.size or B.size != A.size:
        return float("nan")
    out = (ctypes.c_double * 4)()
    _check(lib.matinv_host_defect_f64(A.ctypes.data, B.ctypes.data, n, out))
    return float(np.sqrt(n) - np.sqrt(out[3]))


def invert_f64_dev(A, X=None, piv=None, nopivot: bool = False, flags: int = 0):
    """Device-resident FP64 inversion on torch's current stream; A: (n,n) float64 CUDA tensor."""
    import torch

    assert A.is_cuda and A.dtype == torch.float64 and A.is_contiguous() and A.dim() == 2 and A.shape[0] == A.shape[1]
    if X is None:
        X = torch.empty_like(A)
    st = torch.cuda.current_stream(A.device).cuda_stream
    f = flags | (FLAG_NOPIVOT if nopivot else 0)
    with torch.cuda.device(A.device):
        rc = lib.matinv_invert_f64_dev(_torch_ptr(A), A.shape[0], _torch_ptr(X),
                                       _torch_ptr(piv) if piv is not None else None, ctypes.c_void_p(st), f)
    _check(rc)
    return rc, X


def _torch_ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def invert_dev(A, X=None, piv=None, flags: int = 0):
    """Device-resident inversion on torch's current stream.  A: (n,n) float32 CUDA tensor (contiguous);
    X: output tensor (may be A); piv: optional int32 CUDA tensor of n entries.  Returns the status
    (OK / SINGULAR)."""
    import torch

    assert A.is_cuda and A.dtype == torch.float32 and A.is_contiguous() and A.dim() == 2 and A.shape[0] == A.shape[1]
    if X is None:
        X = torch.empty_like(A)
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        rc = lib.matinv_invert_f32_dev(_torch_ptr(A), A.shape[0], _torch_ptr(X),
                                       _torch_ptr(piv) if piv is not None else None, ctypes.c_void_p(st), flags)
    _check(rc)
    return rc, X


def invert_batched(A: np.ndarray, flags: int = 0):
    """Batched small-n inversion from host memory: A (batch, n, n) -> (X, info)."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    b, n, n2 = A.shape
    assert n == n2
    X = np.empty_like(A)
    info = np.empty(b, dtype=np.int32)
    _check(lib.matinv_invert_batched_f32(A.ctypes.data, n, b, X.ctypes.data, info.ctypes.data, flags))
    return X, info


def invert_sharded(A: np.ndarray, ngpu: int = 0, flags: int = 0, want_piv: bool = False):
    """One inversion column-sharded over `ngpu` GPUs of this process (matinv_invert_sharded_f32; ngpu = 0: MATINV_NGPU or
    every visible device).  Same return convention as invert(): X or None when singular, bit-identical to invert()."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    n = A.shape[0]
    assert A.shape == (n, n)
    X = np.empty_like(A)
    piv = np.empty(n, dtype=np.int32)
    rc = _check(lib.matinv_invert_sharded_f32(A.ctypes.data, n, X.ctypes.data, piv.ctypes.data, ngpu, 0, flags))
    Xr = X if rc == OK else None
    return (Xr, piv) if want_piv else Xr


def sharded_synthetic(n: int, seed: int, kind: str = "uniform", ngpu: int = 0):
    """Column-sharded inversion of the synthetic workload generated on the devices: (rc, piv, compute_ms)."""
    piv = np.empty(n, dtype=np.int32)
    ms = ctypes.c_double(-1.0)
    rc = _check(lib.matinv_sharded_synthetic_f32(n, seed, 1 if kind == "diagdom" else 0, ngpu, piv.ctypes.data, ctypes.byref(ms)))
    return rc, piv, ms.value


def invert_batched_ngpu(A: np.ndarray, ngpu: int = 0, flags: int = 0):
    """Batched small-n inversion split by matrix index over `ngpu` GPUs (no communication): (X, info)."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    b, n, n2 = A.shape
    assert n == n2
    X = np.empty_like(A)
    info = np.empty(b, dtype=np.int32)
    _check(lib.matinv_invert_batched_f32_ngpu(A.ctypes.data, n, b, X.ctypes.data, info.ctypes.data, ngpu, flags))
    return X, info


def invert_batched_dev(A, X=None, info=None, flags: int = 0):
    """Batched small-n inversion on torch's current stream (asynchronous)."""
    import torch

    assert A.is_cuda and A.dtype == torch.float32 and A.is_contiguous() and A.dim() == 3 and A.shape[1] == A.shape[2]
    if X is None:
        X = torch.empty_like(A)
    if info is None:
        info = torch.empty(A.shape[0], dtype=torch.int32, device=A.device)
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        _check(lib.matinv_invert_batched_f32_dev(_torch_ptr(A), A.shape[1], A.shape[0], _torch_ptr(X), _torch_ptr(info),
                                                 ctypes.c_void_p(st), flags))
    return X, info


def generate_dev(n: int, seed: int, kind: str = "uniform", device="cuda"):
    """Synthetic n x n workload on the device (same bits as oracle.gj_oracle.generate)."""
    import torch

    A = torch.empty((n, n), dtype=torch.float32, device=device)
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        _check(lib.matinv_generate_f32_dev(_torch_ptr(A), n, n, seed, 1 if kind == "diagdom" else 0, 0, n,
                                           ctypes.c_void_p(st)))
    return A


def generate_batched_dev(n: int, first: int, count: int, seed0: int, device="cuda"):
    import torch

    A = torch.empty((count, n, n), dtype=torch.float32, device=device)
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        _check(lib.matinv_generate_batched_f32_dev(_torch_ptr(A), n, first, count, seed0, ctypes.c_void_p(st)))
    return A


def residual_dev(A, X):
    """(||A X - I||_F / (n ||A||_F ||X||_F), sqrt(n) - ||A X||_F ~ defect) computed in FP64 on the device."""
    import math

    import torch

    out = (ctypes.c_double * 3)()
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        _check(lib.matinv_residual_f32_dev(_torch_ptr(A), _torch_ptr(X), A.shape[0], out, ctypes.c_void_p(st)))
    n = A.shape[0]
    r2, a2, x2 = out[0], out[1], out[2]
    return math.sqrt(r2) / (n * math.sqrt(a2) * math.sqrt(x2)), r2


def ffma_peak_tflops() -> float:
    import torch

    v = ctypes.c_double()
    st = torch.cuda.current_stream().cuda_stream
    _check(lib.matinv_ffma_peak_tflops(ctypes.byref(v), ctypes.c_void_p(st)))
    return v.value


def probe_residual_dev(A, X):
    """O(n^2) randomised estimate of ||A X - I||_F / (n ||A||_F ||X||_F) on CUDA tensors (the gate of FLAG_TF32X3)."""
    import torch

    n = A.shape[0]
    out = (ctypes.c_double * 3)()
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        _check(lib.matinv_probe_residual_f32_dev(_torch_ptr(A), _torch_ptr(X), n, out, ctypes.c_void_p(st)))
    r2, a2, x2 = out[0], out[1], out[2]
    return float(np.sqrt(r2) / (n * np.sqrt(a2) * np.sqrt(x2)))


def tf32x3_gate_dev(A, X):
    """The acceptance rule of FLAG_TF32X3 applied to a given pair of CUDA tensors: (accepted, estimate, estimate * sqrt(n))."""
    import torch

    n = A.shape[0]
    out = (ctypes.c_double * 2)()
    st = torch.cuda.current_stream(A.device).cuda_stream
    with torch.cuda.device(A.device):
        rc = _check(lib.matinv_tf32x3_gate_dev(_torch_ptr(A), _torch_ptr(X), n, out, ctypes.c_void_p(st)))
    return bool(rc), out[0], out[1]


def tf32x3_status():
    """Bookkeeping of the residual-gated 3xTF32 path (FLAG_TF32X3): dict(estimate, fell_back, inversions, fallbacks).
    `estimate` is the randomised estimate of ||A X - I||_F / (n ||A||_F ||X||_F) of the last gated inversion."""
    est, fb = ctypes.c_double(), ctypes.c_int()
    ninv, nfb = ctypes.c_longlong(), ctypes.c_longlong()
    _check(lib.matinv_tf32x3_status(ctypes.byref(est), ctypes.byref(fb), ctypes.byref(ninv), ctypes.byref(nfb)))
    return {"estimate": est.value, "fell_back": bool(fb.value), "inversions": ninv.value, "fallbacks": nfb.value}


def debug_trailing_update(W, k0: int, CmT, U, mode: int, reps: int = 1):
    """Test hook: one trailing update W[i][j] -= sum_t CmT[t][i] U[t][j] in place on CUDA tensors (W: (npad, npad),
    CmT / U: (128, npad)); mode 0 = FP32 SIMT kernel, 1 = 3xTF32 tcgen05 kernel.  Returns the average ms per application."""
    import torch

    npad = W.shape[0]
    assert W.is_cuda and W.dtype == torch.float32 and W.is_contiguous() and W.shape == (npad, npad)
    assert CmT.shape == (128, npad) and U.shape == (128, npad) and CmT.is_contiguous() and U.is_contiguous()
    ms = ctypes.c_double()
    st = torch.cuda.current_stream(W.device).cuda_stream
    with torch.cuda.device(W.device):
        _check(lib.matinv_debug_trailing_update(_torch_ptr(W), npad, npad, k0, _torch_ptr(CmT), _torch_ptr(U), mode, reps,
                                                ctypes.byref(ms), ctypes.c_void_p(st)))
    return ms.value


def profile_enable(on: bool = True) -> None:
    lib.matinv_profile_enable(1 if on else 0)


def profile_read():
    """dict(gemm_ms, gemm_launches, gemm_flops, launches) accumulated since profile_enable()."""
    ms, fl = ctypes.c_double(), ctypes.c_double()
    gl, al = ctypes.c_longlong(), ctypes.c_longlong()
    _check(lib.matinv_profile_read(ctypes.byref(ms), ctypes.byref(gl), ctypes.byref(fl), ctypes.byref(al)))
    return {"gemm_ms": ms.value, "gemm_launches": gl.value, "gemm_flops": fl.value, "launches": al.value}
