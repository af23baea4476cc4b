"""Tiny driver for ncu captures: python tools/run_single.py N [reps] [batched]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_BATCHED, SEED_UNIFORM

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if len(sys.argv) > 3 and sys.argv[3] == "batched":
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
    A = m.generate_batched_dev(n, 0, batch, SEED_BATCHED)
    X = torch.empty_like(A)
    for _ in range(reps):
        m.invert_batched_dev(A, X)
    torch.cuda.synchronize()
    print("ok batched", n, batch)
else:
    A = m.generate_dev(n, SEED_UNIFORM + n, "uniform")
    X = torch.empty_like(A)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record()
        rc, _ = m.invert_dev(A, X)
        e1.record()
        assert rc == 0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("ok", n, "last ms", round(ms, 2), "TFLOP/s", round(2.0 * n ** 3 / ms / 1e9, 2))
