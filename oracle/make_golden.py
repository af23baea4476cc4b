"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference (oracle/_ref/libref.so, built by
oracle/Makefile.ref from /root/reference) on seeded inputs.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Each fixture stores how the input is regenerated (family, n, modification), the reference's return code
(0 = vector returned, 1 = empty vector) and its output: the full float32 inverse for n = 256, a SHA-256 of the
bytes for larger n.  Two builds of the reference's kernels are recorded, -ffp-contract=off and =fast, because
OpenCL C leaves the contraction of `a - b*c` to the compiler.
"""
import ctypes
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gj_oracle as o  # noqa: E402

GOLD = ROOT / "tests" / "golden"
REF = ROOT / "oracle" / "_ref" / "libref.so"


def make_input(family: str, n: int, mod: str) -> np.ndarray:
    A = {"uniform": o.uniform, "diagdom": o.diagdom, "hollow": lambda k: o.hollow(k)[0]}[family](n)
    if mod == "zero_row7":
        A[7] = 0.0
    elif mod == "all_zero":
        A[:] = 0.0
    elif mod == "nan00":
        A[0, 0] = np.nan
    elif mod == "dup_row":
        A[9] = A[5]
    elif mod == "zero_col11":
        A[:, 11] = 0.0
    elif mod == "zero_diag3":
        A[3, 3] = 0.0
    elif mod != "none":
        raise ValueError(mod)
    return A


def make_input64(family: str, n: int, mod: str) -> np.ndarray:
    """FP64 fixtures: the FP32 workload widened (exactly representable), so the same generators serve both."""
    return make_input(family, n, mod).astype(np.float64)


def run_reference(fn_name: str, A: np.ndarray, contract: str):
    """Runs in a child process: the kernel build mode is read from the environment when the program is built.
    Works for both element types (the bridge takes untyped pointers).  matrix_inversion_no_pivots enqueues its row and
    column kernels over n+1 work-items with a local size of 256 (matrix_inversion_no_pivots.cpp:507): invalid in
    OpenCL 1.2, accepted by drivers that allow a ragged last group -- minicl follows the latter with MINICL_RAGGED=1."""
    code = f"""
import ctypes, numpy as np, sys
L = ctypes.CDLL({str(REF)!r})
f = getattr(L, {fn_name!r}); f.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
A = np.load(sys.argv[1]); n = A.shape[0]; X = np.zeros_like(A)
rc = f(A.ctypes.data, A.size, n, X.ctypes.data)
np.save(sys.argv[2], X); open(sys.argv[3], 'w').write(str(rc))
"""
    tmp = Path("/tmp/minicl_golden"); tmp.mkdir(exist_ok=True)
    np.save(tmp / "a.npy", A)
    env = dict(os.environ, MINICL_FP_CONTRACT=contract, MINICL_CACHE=str(ROOT / "oracle" / "_ref" / "kcache"),
               MINICL_RAGGED="1" if "no_pivots" in fn_name else "0")
    subprocess.run([sys.executable, "-c", code, str(tmp / "a.npy"), str(tmp / "x.npy"), str(tmp / "rc.txt")], env=env, check=True,
                   stdout=subprocess.DEVNULL)
    return int((tmp / "rc.txt").read_text()), np.load(tmp / "x.npy")


FIXTURES = [
    # (name, entry point, family, n, modification, contract)
    ("lib_diagdom256_off", "ref_matrix_inv_32", "diagdom", 256, "none", "off"),
    ("lib_uniform256_off", "ref_matrix_inv_32", "uniform", 256, "none", "off"),
    ("lib_uniform256_fast", "ref_matrix_inv_32", "uniform", 256, "none", "fast"),
    ("lib_hollow256_off", "ref_matrix_inv_32", "hollow", 256, "none", "off"),
    ("lib_uniform512_off", "ref_matrix_inv_32", "uniform", 512, "none", "off"),
    ("lib_diagdom512_fast", "ref_matrix_inv_32", "diagdom", 512, "none", "fast"),
    ("sol_uniform256_off", "ref_matrix_inversion_FP32", "uniform", 256, "none", "off"),
    ("sol_zero_row7", "ref_matrix_inversion_FP32", "uniform", 256, "zero_row7", "off"),
    ("sol_all_zero", "ref_matrix_inversion_FP32", "uniform", 256, "all_zero", "off"),
    ("sol_nan00", "ref_matrix_inversion_FP32", "uniform", 256, "nan00", "off"),
    ("sol_dup_row", "ref_matrix_inversion_FP32", "uniform", 256, "dup_row", "off"),
    ("sol_zero_col11", "ref_matrix_inversion_FP32", "uniform", 256, "zero_col11", "off"),
]

# FP64 entry points of the development copy; outputs stored as SHA-256 only
FIXTURES64 = [
    ("sol64_uniform256_off", "ref_matrix_inversion_FP64", "uniform", 256, "none", "off"),
    ("sol64_uniform256_fast", "ref_matrix_inversion_FP64", "uniform", 256, "none", "fast"),
    ("sol64_hollow512_off", "ref_matrix_inversion_FP64", "hollow", 512, "none", "off"),
    ("sol64_zero_row7", "ref_matrix_inversion_FP64", "uniform", 256, "zero_row7", "off"),
    ("sol64_nan00", "ref_matrix_inversion_FP64", "uniform", 256, "nan00", "off"),
    ("nopiv_diagdom256_off", "ref_matrix_inversion_no_pivots", "diagdom", 256, "none", "off"),
    ("nopiv_diagdom512_fast", "ref_matrix_inversion_no_pivots", "diagdom", 512, "none", "fast"),
    ("nopiv_uniform256_off", "ref_matrix_inversion_no_pivots", "uniform", 256, "none", "off"),
    ("nopiv_zero_diag3", "ref_matrix_inversion_no_pivots", "diagdom", 256, "zero_diag3", "off"),
]

if __name__ == "__main__":
    assert REF.exists(), "build the reference first: make -C oracle -f Makefile.ref"
    GOLD.mkdir(parents=True, exist_ok=True)
    for name, fn, family, n, mod, contract in FIXTURES:
        A = make_input(family, n, mod)
        rc, X = run_reference(fn, A, contract)
        sha = hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() if rc == 0 else ""
        store = X if (rc == 0 and n <= 256 and mod == "none") else np.zeros((0, 0), np.float32)
        np.savez_compressed(GOLD / f"ref_{name}.npz", fn=fn, family=family, n=n, mod=mod, contract=contract, rc=rc, sha256=sha,
                            X=store, finite=bool(np.isfinite(X).all()) if rc == 0 else False)
        print(f"{name:24s} rc={rc} finite={np.isfinite(X).all() if rc == 0 else '-'} sha={sha[:16]}")
    for name, fn, family, n, mod, contract in FIXTURES64:
        A = make_input64(family, n, mod)
        rc, X = run_reference(fn, A, contract)
        finite = bool(np.isfinite(X).all()) if rc == 0 else False
        sha = hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest() if rc == 0 else ""
        np.savez_compressed(GOLD / f"ref_{name}.npz", fn=fn, family=family, n=n, mod=mod, contract=contract, rc=rc, sha256=sha,
                            finite=finite)
        print(f"{name:24s} rc={rc} finite={finite if rc == 0 else '-'} sha={sha[:16]}")
