"""Single-call multi-GPU entries of the C-ABI (csrc/gj_multi.cu): one process, one host thread per GPU, NCCL.

  matinv_invert_sharded_f32        column-sharded inversion == single-GPU inversion, bit for bit (SURVEY.md 8(d) parity gate
                                   "sharded == single-GPU bitwise"), on every GPU count the box offers (1 runs the same
                                   schedule without NCCL, so the entry is covered on a one-GPU box too)
  matinv_invert_batched_f32_ngpu   index split of the batched path == the one-GPU batched result
  matrix_inv_32 with MATINV_NGPU   the C++ surface reaches the sharded entry (compiled caller)
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gj_oracle as o  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


@pytest.fixture(scope="module")
def m():
    import gpu_matrix_inversion_b200 as mod

    return mod


def gpu_counts(m):
    nd = m.device_count()
    return [g for g in (1, 2, 3, 4, 8) if g <= nd]


@pytest.mark.parametrize("n", [300, 1000, 2176])
def test_sharded_entry_equals_single_gpu_and_oracle(m, n):
    A = o.uniform(n)
    Xs, ps = m.invert(A, want_piv=True)
    Xo, po, io = o.invert_inplace(A)
    assert io == 0 and np.array_equal(ps, po) and np.array_equal(bits(Xs), bits(Xo))
    for g in gpu_counts(m):
        X, piv = m.invert_sharded(A, ngpu=g, want_piv=True)
        assert X is not None, (n, g)
        assert np.array_equal(piv, ps), (n, g)
        assert np.array_equal(bits(X), bits(Xs)), (n, g)


def test_sharded_entry_singular_and_argument_checks(m):
    n = 640
    S = o.uniform(n)
    S[7] = 0.0
    for g in gpu_counts(m):
        assert m.invert_sharded(S, ngpu=g) is None
    with pytest.raises(m.MatinvError):
        m.invert_sharded(o.uniform(256), ngpu=m.device_count() + 1)
    A = o.uniform(256)
    X = np.empty_like(A)
    assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 256, X.ctypes.data, None, 1, 64, 0) == m.E_UNSUPPORTED  # nb must be 0 / 128
    assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 0, X.ctypes.data, None, 1, 0, 0) == m.E_INVALID
    if m.device_count() > 1:
        assert m.lib.matinv_nccl_version() >= 20000


def test_sharded_synthetic_matches_generator(m):
    """The device-generated workload of the timing entry is the oracle's matrix: same pivot sequence as the host path."""
    n = 1500
    for g in gpu_counts(m):
        rc, piv, ms = m.sharded_synthetic(n, o.SEED_UNIFORM + n, "uniform", ngpu=g)
        assert rc == 0 and ms > 0
        assert np.array_equal(piv, o.invert_inplace(o.uniform(n))[1]), g


def test_batched_index_split(m):
    A = o.batched(64, 0, 301)
    A[17] = 0.0
    X1, i1 = m.invert_batched(A)
    for g in gpu_counts(m):
        Xg, ig = m.invert_batched_ngpu(A, ngpu=g)
        assert np.array_equal(ig != 0, i1 != 0), g
        ok = i1 == 0
        assert np.array_equal(bits(Xg[ok]), bits(X1[ok])), g


def test_cpp_caller_reaches_the_sharded_entry(m, tmp_path):
    """A caller compiled against include/mat_inv_32.h with MATINV_NGPU set: same bits as without it."""
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include "mat_inv_32.h"
#include <cstdio>
#include <cstdint>
#include <cstring>
int main(int argc, char **argv) {
    const int n = 4224;
    std::vector<float> a((size_t)n * n);
    uint64_t s = 12345;
    for (auto &x : a) { s = s * 6364136223846793005ull + 1442695040888963407ull; x = (float)((s >> 40) % 1000) / 10.0f; }
    std::vector<float> r = matrix_inv_32(a, n);
    if (r.size() != a.size()) { std::printf("EMPTY\n"); return 1; }
    uint64_t h = 1469598103934665603ull;
    for (float v : r) { uint32_t b; std::memcpy(&b, &v, 4); h = (h ^ b) * 1099511628211ull; }
    std::printf("HASH %016llx\n", (unsigned long long)h);
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = ROOT / "gpu_matrix_inversion_b200"
    subprocess.run(["g++", "-O2", "-std=c++14", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(libdir),
                    "-lmatinv32", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    outs = []
    for ng in [None] + [g for g in gpu_counts(m) if g > 1][:2] + [1]:
        env = {k: v for k, v in os.environ.items() if k != "MATINV_NGPU"}
        if ng is not None:
            env["MATINV_NGPU"] = str(ng)
            env["MATINV_NGPU_MIN_ORDER"] = "4096"      # the default threshold (32768) would keep this order on one GPU
        p = subprocess.run([str(exe)], env=env, capture_output=True, text=True, timeout=600)
        lines = [l for l in p.stdout.splitlines() if l.startswith("HASH")]     # (NCCL may print its version banner first)
        assert p.returncode == 0 and len(lines) == 1, (ng, p.stdout, p.stderr[-2000:])
        outs.append(lines[0])
    assert len(set(outs)) == 1, outs
