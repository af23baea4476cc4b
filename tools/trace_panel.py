"""Print the in-kernel timeline (SM cycles) of the last sub-panel / panel-update launches."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import gpu_matrix_inversion_b200 as m
from oracle.gj_oracle import SEED_UNIFORM
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
A = m.generate_dev(n, SEED_UNIFORM + n, "uniform"); X = torch.empty_like(A)
m.invert_dev(A, X)
m.lib.matinv_debug_trace(1, None)
m.invert_dev(A, X)
buf = (ctypes.c_longlong * 128)()
m.lib.matinv_debug_trace(0, buf)
t = list(buf)
print("subpanel: load", t[1]-t[0], "csync", t[2]-t[1], "steps", [t[3+i]-t[2+i] for i in range(16)], "wb", t[20]-t[18], "end", t[21]-t[20], "total", t[21]-t[0])
print("update cta3: prologue", t[33]-t[32], "gather", t[34]-t[33], "recur", t[35]-t[34], "main", t[36]-t[35], "total", t[36]-t[32])
print("update bookkeeping: ps", t[49]-t[48], "hist", t[50]-t[49])
print("step 8 detail: cand+redux", t[65]-t[64], "sync1", t[66]-t[65], "warp0 push", t[67]-t[66], "cluster barrier", t[68]-t[67], "warp0 reduce+div", t[69]-t[68], "sync2", t[70]-t[69], "update", t[11]-t[70])
