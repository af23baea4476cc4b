"""Per-phase cycle accounting of the packed-pair batched kernel (warp 0 of CTA 0, first matrix; SM cycles)."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gpu_matrix_inversion_b200 as m
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 2 * 4 * 4
A = m.generate_batched_dev(64, 0, batch, 0xB2002000)
X = torch.empty_like(A)
m.invert_batched_dev(A, X)
m.lib.matinv_debug_trace(1, None)
m.invert_batched_dev(A, X)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
m.lib.matinv_debug_trace(0, buf)
t = list(buf)[96:104]
names = ["search", "publish+sync", "divide+sync", "update", "reload", "bookkeeping", "rotate/end"]
tot = sum(t[:7])
print("batch", batch, "total cycles for one matrix", tot, "per step", tot / 64)
for n_, v in zip(names, t):
    print(f"  {n_:14s} {v:8d}  {v / 64:7.1f} per step  {100 * v / max(tot, 1):5.1f} %")
