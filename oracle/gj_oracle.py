"""ctypes front-end of oracle/gj_oracle.c plus a tiny pure-numpy restatement.

TEST INFRASTRUCTURE ONLY -- see the header of gj_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by
gpu_matrix_inversion_b200 (the product path has no CPU fallback).

Reference being restated: /root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:11-395
(kernels :12-204, step loop :317-362) and the singular check of
matrix_inv_solution/.../matrix_inversion_FP32.cpp:814-835.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libgj_oracle.so"
_SRC = _HERE / "gj_oracle.c"

NOFMA = 1
QUIRK = 2  # gj_aug_f32 only: the reference's pivot search as written (n % 256 == 0)
NOPIVOT = 4  # pivot = diagonal entry, no row interchange (matrix_inversion_no_pivots.cpp)

SEED_UNIFORM = 0xB2000000
SEED_DIAGDOM = 0xB2001000
SEED_BATCHED = 0xB2002000


def build(force: bool = False) -> Path:
    """gcc the C restatement in-tree (the .so is git-ignored but travels with gpurun)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < _SRC.stat().st_mtime:
        cmd = ["gcc", "-O3", "-mfma", "-mavx2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
               "-o", str(_SO), str(_SRC), "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        L.gj_generate_f32.argtypes = [fp, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        L.gj_generate_f32.restype = None
        L.gj_generate_hollow_f32.argtypes = [fp, ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        L.gj_generate_hollow_f32.restype = None
        for suf, p in (("f32", fp), ("f64", dp)):
            getattr(L, f"gj_aug_{suf}").argtypes = [p, ctypes.c_int, p, ip, ip, ctypes.c_int]
            getattr(L, f"gj_inplace_{suf}").argtypes = [p, ctypes.c_int, p, ip, ctypes.c_int]
            getattr(L, f"gj_blocked_{suf}").argtypes = [p, ctypes.c_int, p, ip, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_int]
            for f in ("aug", "inplace", "blocked"):
                getattr(L, f"gj_{f}_{suf}").restype = ctypes.c_int
        L.gj_residual_f32.argtypes = [fp, fp, ctypes.c_int, dp]
        L.gj_residual_f32.restype = ctypes.c_double
        L.gj_oracle_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _ct(dtype):
    return ctypes.c_float if dtype == np.float32 else ctypes.c_double


# ------------------------------------------------------------------ generators (SURVEY s.8d)

def generate(n: int, seed: int, kind: str = "uniform") -> np.ndarray:
    """Counter-based U[0,100) FP32 matrix; kind in {'uniform', 'diagdom'}."""
    A = np.empty((n, n), dtype=np.float32)
    lib().gj_generate_f32(_ptr(A, ctypes.c_float), n, ctypes.c_uint64(seed), 1 if kind == "diagdom" else 0)
    return A


def uniform(n: int) -> np.ndarray:
    return generate(n, SEED_UNIFORM + n, "uniform")


def diagdom(n: int) -> np.ndarray:
    return generate(n, SEED_DIAGDOM + n, "diagdom")


def batched(n: int, first: int, count: int) -> np.ndarray:
    """Matrices first..first+count-1 of the batched workload (matrix b uses seed SEED_BATCHED + b)."""
    out = np.empty((count, n, n), dtype=np.float32)
    for b in range(count):
        out[b] = generate(n, SEED_BATCHED + first + b, "uniform")
    return out


def hollow(n: int, state: int = 1):
    """Hollow rand()%10 matrix of SOL/main_file.cpp:41-52 (MSVC LCG); returns (A, next_state)."""
    A = np.empty((n, n), dtype=np.float32)
    st = ctypes.c_uint32(state)
    lib().gj_generate_hollow_f32(_ptr(A, ctypes.c_float), n, ctypes.byref(st))
    return A, st.value


# ------------------------------------------------------------------ inversions

def _run(fn, A: np.ndarray, *extra):
    A = np.ascontiguousarray(A)
    n = A.shape[0]
    assert A.shape == (n, n)
    X = np.empty_like(A)
    piv = np.full(n, -1, dtype=np.int32)
    info = fn(A, n, X, piv, *extra)
    return X, piv, info


def invert_aug(A, flags: int = 0, forced_piv=None):
    """A.1/A.2: explicit [A|I] -- the reference's formulation.  Returns (X, piv, info)."""
    A = np.ascontiguousarray(A)
    ct = _ct(A.dtype)
    suf = "f32" if A.dtype == np.float32 else "f64"
    f = getattr(lib(), f"gj_aug_{suf}")
    fp = None
    if forced_piv is not None:
        forced_piv = np.ascontiguousarray(forced_piv, dtype=np.int32)
        fp = _ptr(forced_piv, ctypes.c_int)
    return _run(lambda a, n, x, p: f(_ptr(a, ct), n, _ptr(x, ct), _ptr(p, ctypes.c_int), fp, flags), A)


def invert_inplace(A, flags: int = 0):
    """A.3: in-place form, what the unblocked kernels compute."""
    A = np.ascontiguousarray(A)
    ct = _ct(A.dtype)
    f = getattr(lib(), "gj_inplace_f32" if A.dtype == np.float32 else "gj_inplace_f64")
    return _run(lambda a, n, x, p: f(_ptr(a, ct), n, _ptr(x, ct), _ptr(p, ctypes.c_int), flags), A)


def invert_blocked(A, nb: int = 32, w: int = 8, flags: int = 0):
    """A.4: blocked right-looking form with W-wide sub-panels, the schedule of the CUDA path."""
    A = np.ascontiguousarray(A)
    ct = _ct(A.dtype)
    f = getattr(lib(), "gj_blocked_f32" if A.dtype == np.float32 else "gj_blocked_f64")
    return _run(lambda a, n, x, p: f(_ptr(a, ct), n, _ptr(x, ct), _ptr(p, ctypes.c_int), nb, w, flags), A)


def residual(A: np.ndarray, X: np.ndarray):
    """(||A X - I||_F / (N ||A||_F ||X||_F), sqrt(N) - ||A X||_F) with FP64 accumulation."""
    A = np.ascontiguousarray(A, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    d = ctypes.c_double(0.0)
    r = lib().gj_residual_f32(_ptr(A, ctypes.c_float), _ptr(X, ctypes.c_float), A.shape[0], ctypes.byref(d))
    return r, d.value


def threads() -> int:
    return int(lib().gj_oracle_threads())


# ------------------------------------------------------------------ pure-numpy restatement (small n)

def invert_numpy_f32(A: np.ndarray):
    """Independent float32 restatement of A.1 for small n (no FMA: numpy rounds the product).

    Used only to cross-check gj_aug_f32(flags=NOFMA) -- two implementations of the same text.
    """
    n = A.shape[0]
    M = np.zeros((n, 2 * n), dtype=np.float32)
    M[:, :n] = A
    M[:, n:] = np.eye(n, dtype=np.float32)
    piv = np.full(n, -1, dtype=np.int32)
    for r in range(n):
        col = np.abs(M[r:, r])
        p, best = r, col[0]
        for k in range(1, n - r):  # strict '>' upward scan; NaN never wins
            if col[k] > best:
                best, p = col[k], r + k
        v = M[p, r]
        piv[r] = p
        if v == 0 or not np.isfinite(v):
            return None, piv, r + 1
        if p != r:
            M[[r, p]] = M[[p, r]]
        M[r] = M[r] / v
        for i in range(n):
            c = M[i, r]
            if i == r or c == 0:
                continue
            M[i] = M[i] - (c * M[r]).astype(np.float32)
    X = M[:, n:].copy()
    return X, piv, (0 if np.isfinite(X).all() else -1)
