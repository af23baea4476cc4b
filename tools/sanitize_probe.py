"""Small end-to-end pass over every product kernel family with residual checks -- a quick health check of a build, and
the input for `compute-sanitizer --tool memcheck` where that tool is available (it is closed on the round-1 GPU pool).
FP32 blocked path with look-ahead (N=1500), unblocked FP32 (N=200), batched 64x64 and 32x32, FP64 blocked and unblocked
with and without pivoting (N=130, 321), residual kernels."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import gpu_matrix_inversion_b200 as m

rng = np.random.default_rng(7)
A = (rng.random((1500, 1500)) * 100).astype(np.float32)
X = m.invert(A)
assert X is not None and np.abs(A.astype(np.float64) @ X - np.eye(1500)).max() < 1e-1
B = (rng.random((200, 200)) * 100).astype(np.float32)
assert m.invert(B, flags=m.FLAG_UNBLOCKED) is not None
for n in (64, 32):
    Bt = (rng.random((40, n, n)) * 100).astype(np.float32)
    Xb, info = m.invert_batched(Bt)
    assert int((info != 0).sum()) == 0
for n in (130, 321):
    D = rng.random((n, n)) * 100
    for fl in (0, m.FLAG_UNBLOCKED):
        X64 = m.invert_f64(D, flags=fl)
        assert X64 is not None and np.abs(D @ X64 - np.eye(n)).max() < 1e-8
        Xn = m.invert_f64(D + np.eye(n) * 100 * n, nopivot=True, flags=fl)
        assert Xn is not None
    assert abs(m.matrix_multiply(X64.ravel(), D.ravel())) < 1e-8
d = torch.from_numpy(A).cuda()
rc, Xd = m.invert_dev(d)
r, _ = m.residual_dev(d, Xd)
assert rc == 0 and r < 1e-5
torch.cuda.synchronize()
print("sanitize probe ok")
