"""Golden SHA-256 of the oracle's (pivot sequence || inverse) for the full-size single-GPU workloads.

    python oracle/make_golden_large.py [orders...]        (default: 8320 16384)

The blocked OpenMP oracle (gj_blocked_f32: the A.4 schedule, proven bit-identical to the unblocked A.3 / A.1 forms by
tests/test_oracle.py) replays the synthetic uniform workload of the named order -- the matrix bench.py inverts on rank 0
(seed SEED_UNIFORM + n) -- and records sha256(piv as int32 || X as float32).  tests/test_gpu_parity.py::
test_full_size_golden_hash asserts that the GPU path reproduces the hash, i.e. pivots AND inverse bit for bit at
BASELINE config 3's order.  N=16384 takes a few minutes on 8 host cores; the result is committed under tests/golden/.
Test infrastructure only.
"""
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gj_oracle as o  # noqa: E402

OUT = ROOT / "tests" / "golden" / "large_sha256.json"


def main():
    orders = [int(a) for a in sys.argv[1:]] or [8320, 16384]
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for n in orders:
        A = o.generate(n, o.SEED_UNIFORM + n, "uniform")
        t0 = time.time()
        X, piv, info = o.invert_blocked(A, nb=128, w=16)
        dt = time.time() - t0
        assert info == 0
        h = hashlib.sha256(np.asarray(piv, dtype=np.int32).tobytes() + np.ascontiguousarray(X, dtype=np.float32).tobytes()).hexdigest()
        res, _ = o.residual(A, X) if n <= 4096 else (None, None)
        data[str(n)] = {"n": n, "kind": "uniform", "seed": int(o.SEED_UNIFORM + n), "sha256_piv_X": h,
                        "piv_head": [int(p) for p in piv[:8]], "swaps": int((np.asarray(piv) != np.arange(n)).sum()),
                        "oracle": "gj_blocked_f32(nb=128, w=16)", "seconds": round(dt, 1), "threads": o.threads()}
        print(n, h, f"{dt:.1f}s", flush=True)
        OUT.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
