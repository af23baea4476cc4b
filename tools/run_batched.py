"""Tiny driver for profiling the batched kernel: python tools/run_batched.py [n] [batch] [reps]  (MATINV_BATCHED selects the kernel)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gpu_matrix_inversion_b200 as m  # noqa: E402
from oracle.gj_oracle import SEED_BATCHED  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
A = m.generate_batched_dev(n, 0, batch, SEED_BATCHED)
X = torch.empty_like(A)
info = torch.empty(batch, dtype=torch.int32, device="cuda")
for _ in range(reps):
    m.invert_batched_dev(A, X, info)
torch.cuda.synchronize()
assert int((info != 0).sum()) == 0
print("ok", n, batch, reps)
