"""A/B check of the batched kernels: python tools/batched_ab.py [modes...]   (default: 3 4)

Each mode (MATINV_BATCHED value, read once per process) runs in its own subprocess on the same inputs: 2^18 synthetic
64x64 matrices + the edge-case set of tests/test_gpu_parity.py::test_batched_edge_cases.  Reports inversions/s (CUDA
events, 5 timed launches after 2 warm-ups) and a SHA-256 of (info, X of every non-singular matrix); the hashes of all
modes must agree (the kernels are bit-identical by construction) and mode 3 is checked against the oracle by the test
suite.  Test / tuning aid only.
"""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import hashlib, json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
import gpu_matrix_inversion_b200 as m
from oracle import gj_oracle as o
n = int(sys.argv[1]); batch = int(sys.argv[2])
A = m.generate_batched_dev(n, 0, batch, o.SEED_BATCHED)
X = torch.empty_like(A); info = torch.empty(batch, dtype=torch.int32, device="cuda")
for _ in range(2): m.invert_batched_dev(A, X, info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K): m.invert_batched_dev(A, X, info)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
h = hashlib.sha256()
h.update(info.cpu().numpy().tobytes()); h.update(X[:8192].cpu().numpy().tobytes())
# edge cases (same construction as the parity test)
rng = np.random.default_rng(1234 + n)
mats = [rng.integers(-2, 3, size=(n, n)).astype(np.float32) for _ in range(12)]
for _ in range(4):
    P = np.eye(n, dtype=np.float32)[rng.permutation(n)]
    mats.append(P * rng.choice([-1.0, 1.0], size=(n, 1)).astype(np.float32))
mats.append(np.eye(n, dtype=np.float32)[::-1].copy())
H = o.batched(n, 100, 4)
H[0][rng.random((n, n)) < 0.7] = 0.0
H[1][np.arange(n), np.arange(n)] = 0.0
H[2] *= np.float32(1e-41)
H[3][:, 3] = -H[3][:, 3]; H[3][5, :] = -0.0; H[3][5, 5] = 1.0
mats.extend(H)
bad = o.batched(n, 200, 3)
bad[0][n // 2, n // 3] = np.nan; bad[1][1, 1] = np.inf; bad[2][0, 0] = np.nan
mats.extend(bad)
E = np.ascontiguousarray(np.stack(mats), dtype=np.float32)
Xe, ie = m.invert_batched(E)
nz = (ie != 0)
h.update(nz.tobytes())
for b in range(E.shape[0]):
    if not nz[b]: h.update(Xe[b].tobytes())
oracle_ok = None
if len(sys.argv) > 3 and sys.argv[3] == "oracle":
    oracle_ok = True
    for b in range(E.shape[0]):
        Xo, _, io = o.invert_inplace(E[b])
        if (io != 0) != bool(nz[b]) or (io == 0 and not np.array_equal(Xo.view(np.uint32), Xe[b].view(np.uint32))):
            oracle_ok = False; print("edge case", b, "differs from the oracle", file=sys.stderr)
    Xh = X[:64].cpu().numpy()
    for b in range(64):
        Xo, _, io = o.invert_inplace(o.batched(n, b, 1)[0])
        if io != 0 or not np.array_equal(Xo.view(np.uint32), Xh[b].view(np.uint32)):
            oracle_ok = False; print("matrix", b, "differs from the oracle", file=sys.stderr)
print(json.dumps({"n": n, "batch": batch, "ms": ms, "inv_per_s": batch / (ms * 1e-3), "singular": int(nz.sum()),
                  "sha256": h.hexdigest(), "oracle_ok": oracle_ok}))
"""


def main():
    modes = sys.argv[1:] or ["3", "4"]
    n = int(os.environ.get("AB_N", "64"))
    batch = int(os.environ.get("AB_BATCH", str(1 << 18)))
    out = {}
    for mode in modes:
        env = dict(os.environ, MATINV_BATCHED=mode.split(":")[0])
        for kv in mode.split(":")[1:]:
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": str(ROOT)}, str(n), str(batch), "oracle"], env=env,
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(f"mode {mode}: FAILED\n{r.stdout}\n{r.stderr[-3000:]}")
            out[mode] = None
            continue
        res = json.loads(r.stdout.strip().splitlines()[-1])
        out[mode] = res
        print(f"mode {mode}: {res['inv_per_s']:.3e} inv/s  ({res['ms']:.2f} ms)  oracle_ok={res['oracle_ok']}  sha={res['sha256'][:16]}"
              + (f"\n{r.stderr[-1500:]}" if r.stderr.strip() else ""))
    hs = {v["sha256"] for v in out.values() if v}
    print("hashes agree" if len(hs) == 1 else "HASH MISMATCH")
    sys.exit(0 if len(hs) == 1 and all(v and v["oracle_ok"] for v in out.values()) else 1)


if __name__ == "__main__":
    main()
