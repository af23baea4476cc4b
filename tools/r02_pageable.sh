#!/bin/bash
O=gpurun_out/r02_pageable.txt; : > $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined_upload" >> $O 2>&1; tail -2 $O
for cfg in "MATINV_STAGING=0" "MATINV_STAGING=1"; do
env $cfg python - >> $O 2>&1 <<PY
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
for n in (16384, 8192, 4096):
    A = m.generate_dev(n, SEED_UNIFORM + n, "uniform").cpu().numpy()       # pageable host memory, like std::vector
    X = np.empty_like(A)
    for _ in range(2):
        assert m.lib.matinv_invert_f32(A.ctypes.data, n, X.ctypes.data, None, 0) == 0
    t0 = time.perf_counter(); K = 5
    for _ in range(K):
        rc = m.lib.matinv_invert_f32(A.ctypes.data, n, X.ctypes.data, None, 0)
    dt = (time.perf_counter() - t0) / K
    print("$cfg pageable n=%d e2e %.2f ms = %.1f TFLOP/s" % (n, dt * 1e3, 2.0 * n ** 3 / dt / 1e12), {k: round(v * 1e3, 2) for k, v in m.last_phases().items()}, flush=True)
PY
MATINV_VERBOSE=1 env $cfg tools/time_matrix_inv_32 16384 3 >> $O 2>&1
done
grep -E "e2e|matrix_inv_32|passed|failed" $O
