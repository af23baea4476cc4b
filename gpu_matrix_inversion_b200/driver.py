"""Driver script mirroring /root/reference/matrix_inv_pyopencl.py with the B200 backend.

Same entry points and report format as the reference driver:

* ``matrix_inv(file, N)`` (matrix_inv_pyopencl.py:15-352): draw an ``N x N`` ``U(0,100)`` float32 matrix
  (:17), invert it on the GPU, compute ``err = sqrt(N) - sqrt(sum((X @ A) @ (X @ A)))`` (:341-345) and append
  ``"N t_compute t_total err"`` to ``file`` (:352).  ``t_compute`` / ``t_total`` are the shim's
  "Tempo Computazione" / "Tempo Totale" windows (matinv_last_timing).
* ``__main__``: the sweep N = 10, 20, ..., 2000, 3000, ..., 15000 (:358-371).

Only the backend import changed: PyOpenCL context/queue/program/kernel calls (:24-321) are replaced by one
call into the C-ABI (gpu_matrix_inversion_b200.invert -> matinv_invert_f32).
"""
from __future__ import annotations

import math
import sys

import numpy as np

import gpu_matrix_inversion_b200 as backend

REP = 1  # matrix_inv_pyopencl.py:13


def matrix_inv(file, N, rng=None):
    rng = rng or np.random.default_rng()
    matrice_input = rng.uniform(0, 100, (N, N)).astype(np.float32)
    matrice_input2 = matrice_input.copy()

    inversa = backend.invert(matrice_input)
    if inversa is None:  # singular draw: the library returns an empty vector; the reference script has no such path
        file.write(f"{N} nan nan nan\n")
        return None
    t_total, t_compute = backend.last_timing()

    # CONTROLLO FINALE (matrix_inv_pyopencl.py:341-345)
    c = np.matmul(inversa, matrice_input2)
    vec = c @ c
    somma = float(np.sum(vec))
    errore = math.sqrt(N) - math.sqrt(somma) if somma >= 0 else float("nan")
    print(f"errore: {errore}")
    file.write(f"{N} {t_compute} {t_total} {errore}\n")
    return errore


def sweep(path: str, limit: int = 16000):
    with open(path, "w") as file:
        i = 10
        while i < limit:
            matrix_inv(file, i)
            if i < 2000:
                i += 10
            else:
                i += 1000
    print("fine")


if __name__ == "__main__":
    sweep(sys.argv[1] if len(sys.argv) > 1 else "b200_32.txt", int(sys.argv[2]) if len(sys.argv) > 2 else 16000)
