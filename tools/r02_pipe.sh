#!/bin/bash
O=gpurun_out/r02_pipe.txt; : > $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined_upload or phase_timers or golden" >> $O 2>&1; tail -3 $O
for cfg in "MATINV_H2D_PIPELINE=0" "MATINV_H2D_CRITICAL=0" "MATINV_H2D_CRITICAL=1" "MATINV_H2D_WINDOWS=4"; do
env $cfg python - >> $O 2>&1 <<PY
import os, sys, time
sys.path.insert(0, ".")
import torch
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
for n in (16384, 8192):
    A = m.generate_dev(n, SEED_UNIFORM + n, "uniform")
    Ah = torch.empty((n, n), dtype=torch.float32, pin_memory=True); Xh = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    Ah.copy_(A); torch.cuda.synchronize()
    for _ in range(2):
        assert m.lib.matinv_invert_f32(Ah.data_ptr(), n, Xh.data_ptr(), None, 0) == 0
    t0 = time.perf_counter(); K = 8
    for _ in range(K):
        rc = m.lib.matinv_invert_f32(Ah.data_ptr(), n, Xh.data_ptr(), None, 0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print("$cfg", "n=%d e2e %.2f ms = %.1f TFLOP/s" % (n, dt * 1e3, 2.0 * n ** 3 / dt / 1e12), "phases", {k: round(v * 1e3, 2) for k, v in m.last_phases().items()}, flush=True)
PY
done
grep "e2e" $O
