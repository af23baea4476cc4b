"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, the C++ symbol has the reference's mangled name, and the argument conventions of
matrix_inv_32 (mat_inv_32.cpp:207-215) hold without a GPU."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import gpu_matrix_inversion_b200 as m

    header = (ROOT / "include" / "matinv_shim.h").read_text()
    declared = sorted(set(re.findall(r"\b(matinv_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "gpu_matrix_inversion_b200" / "libmatinv32.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    missing = [d for d in declared if d not in exported]
    assert not missing, missing
    assert sorted(m.EXPORTS) == declared
    # the C++ surface: Itanium mangling of std::vector<float> matrix_inv_32(std::vector<float>, int)
    assert "_Z13matrix_inv_32St6vectorIfSaIfEEi" in exported
    # the development copy's function set (include/matrix_inversion.h <- SOL/headers.h:5-11)
    for sym in ("_Z21matrix_inversion_FP32St6vectorIfSaIfEEi", "_Z21matrix_inversion_FP64St6vectorIdSaIdEEi",
                "_Z26matrix_inversion_no_pivotsSt6vectorIdSaIdEEi", "_Z15matrix_multiplySt6vectorIdSaIdEES1_"):
        assert sym in exported, sym


def test_python_twin_argument_conventions_without_device():
    import gpu_matrix_inversion_b200 as m

    assert m.matrix_inv_32(np.ones(4, np.float32), 0).size == 0        # N <= 0
    assert m.matrix_inv_32(np.ones(4, np.float32), -3).size == 0
    assert m.matrix_inv_32(np.ones(5, np.float32), 3).size == 0        # int(5/3)=1 != 3
    assert m.matrix_inv_32(np.ones(12, np.float32), 3).size == 0       # int(12/3)=4 != 3
    if m.device_count() == 0:
        # no CPU fallback: a valid request fails loudly at the ABI and yields {} at the library surface
        assert m.matrix_inv_32(np.eye(2, dtype=np.float32).ravel(), 2).size == 0
        with pytest.raises(m.MatinvError) as e:
            m.invert(np.eye(2, dtype=np.float32))
        assert e.value.code == m.E_NODEVICE
        with pytest.raises(m.MatinvError):
            m.invert_batched(np.zeros((1, 4, 4), np.float32))


def test_multi_gpu_entries_argument_checks_without_device():
    """The single-call multi-GPU entries (csrc/gj_multi.cu) validate their arguments before touching a device and, like every
    compute entry, fail loudly without one (no CPU fallback)."""
    import gpu_matrix_inversion_b200 as m

    A = np.eye(4, dtype=np.float32)
    X = np.empty_like(A)
    assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 0, X.ctypes.data, None, 1, 0, 0) == m.E_INVALID
    assert m.lib.matinv_invert_sharded_f32(None, 4, X.ctypes.data, None, 1, 0, 0) == m.E_INVALID
    assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 4, X.ctypes.data, None, 1, 64, 0) == m.E_UNSUPPORTED      # nb must be 0 / 128
    assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 4, X.ctypes.data, None, 1, 0, m.FLAG_TF32X3) == m.E_UNSUPPORTED
    B = np.zeros((2, 4, 4), np.float32)
    assert m.lib.matinv_invert_batched_f32_ngpu(B.ctypes.data, 200, 2, B.ctypes.data, None, 1, 0) == m.E_INVALID       # n > 128
    assert m.lib.matinv_invert_batched_f32_ngpu(B.ctypes.data, 4, 0, B.ctypes.data, None, 1, 0) == m.OK               # empty batch
    if m.device_count() == 0:
        assert m.lib.matinv_invert_sharded_f32(A.ctypes.data, 4, X.ctypes.data, None, 1, 0, 0) == m.E_NODEVICE
        assert m.lib.matinv_invert_batched_f32_ngpu(B.ctypes.data, 4, 2, B.ctypes.data, None, 1, 0) == m.E_NODEVICE
        assert m.lib.matinv_sharded_synthetic_f32(256, 1, 0, 1, None, None) == m.E_NODEVICE
        with pytest.raises(m.MatinvError):
            m.invert_sharded(A, ngpu=1)
        assert m.last_phases() is None


def test_pipelined_upload_plan_invariants():
    """Host logic of the pipelined upload (csrc/matinv_shim.cu:plan_pipeline), no GPU needed: for every order the column
    windows tile [0, npad) in multiples of 128, join in order, and every window is active before its first block becomes the
    look-ahead block of the panel chain -- the invariant schedule_lookahead_pipelined relies on."""
    import ctypes

    import gpu_matrix_inversion_b200 as m

    def plan(n, env_windows=None):
        nwin = ctypes.c_int(0)
        ring = ctypes.c_int(0)
        c0 = (ctypes.c_int * 8)()
        nc = (ctypes.c_int * 8)()
        act = (ctypes.c_int * 8)()
        rc = m.lib.matinv_debug_pipeline_plan(n, ctypes.byref(nwin), c0, nc, act, ctypes.byref(ring))
        return rc, nwin.value, list(c0)[:nwin.value], list(nc)[:nwin.value], list(act)[:nwin.value], ring.value

    assert m.lib.matinv_debug_pipeline_plan(0, None, None, None, None, None) == m.E_INVALID
    assert plan(4096)[0] == 0 and plan(8191)[0] == 0            # small orders are not pipelined
    assert plan(70000)[0] == 0                                  # beyond the cluster panel path (n > 65536)
    orders = list(range(8192, 20000, 257)) + [8192, 8320, 9001, 16384, 16385, 32768, 32896, 40000, 65536]
    for n in orders:
        rc, nwin, c0, nc, act, ring = plan(n)
        assert rc == 1 and 2 <= nwin <= 6, n
        npad = (n + 127) // 128 * 128
        assert c0[0] == 0 and sum(nc) == npad, (n, c0, nc)
        assert all(c % 128 == 0 and w % 128 == 0 and w >= 8 * 128 for c, w in zip(c0, nc)), (n, c0, nc)
        assert all(c0[w] == c0[w - 1] + nc[w - 1] for w in range(1, nwin)), (n, c0, nc)
        assert act[0] == 0 and all(act[w] >= max(1, act[w - 1]) for w in range(1, nwin)), (n, act)
        assert all(act[w] <= c0[w] // 128 - 1 for w in range(1, nwin)), (n, act, c0)
        assert ring == max(act) + 2
        assert c0[-1] < n                                       # the last window holds real columns, not only padding


def test_cpp_surface_links_and_follows_conventions(tmp_path):
    """Compile a caller against include/mat_inv_32.h exactly like a user of the reference header."""
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include "mat_inv_32.h"
#include <cstdio>
int main() {
    std::vector<float> a = {4, 7, 2, 6};
    int bad = 0;
    bad |= !matrix_inv_32(a, 0).empty();
    bad |= !matrix_inv_32(a, -1).empty();
    bad |= !matrix_inv_32(a, 3).empty();               // int(4/3) = 1 != 3
    std::vector<float> b = {4, 7, 2, 6, 99};           // size = N*N + k, k < N: accepted, tail ignored
    std::vector<float> r = matrix_inv_32(b, 2);
    std::vector<float> s = matrix_inv_32(std::vector<float>{1, 2, 2, 4}, 2);   // singular
    std::printf("%d %zu %zu\n", bad, r.size(), s.size());
    if (r.size() == 4) std::printf("%.6f %.6f %.6f %.6f\n", r[0], r[1], r[2], r[3]);
    return bad;
}
''')
    exe = tmp_path / "caller"
    libdir = ROOT / "gpu_matrix_inversion_b200"
    subprocess.run(["g++", "-std=c++14", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(libdir),
                    "-lmatinv32", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    first = p.stdout.splitlines()[0].split()
    assert first[0] == "0"
    import gpu_matrix_inversion_b200 as m

    if m.device_count() == 0:
        assert first[1] == "0" and first[2] == "0"          # no device -> {} (never UB like platforms[0])
    else:
        assert first[1] == "4" and first[2] == "0"          # inverse returned; singular -> {}
        vals = [float(x) for x in p.stdout.splitlines()[1].split()]
        assert np.allclose(vals, [0.6, -0.7, -0.2, 0.4], atol=1e-6)


def test_fp64_surface_links_and_follows_conventions(tmp_path):
    """include/matrix_inversion.h compiled by a caller, like a user of the reference's headers.h."""
    src = tmp_path / "caller64.cpp"
    src.write_text(r'''
#include "matrix_inversion.h"
#include <cmath>
#include <cstdio>
int main() {
    std::vector<double> a = {4, 7, 2, 6};
    int bad = 0;
    bad |= !matrix_inversion_FP64(a, 0).empty();
    bad |= !matrix_inversion_FP64(a, 3).empty();            // int(4/3) = 1 != 3
    bad |= !matrix_inversion_no_pivots(a, -2).empty();
    std::vector<double> r = matrix_inversion_FP64(a, 2);
    std::vector<double> q = matrix_inversion_no_pivots(a, 2);
    std::vector<double> s = matrix_inversion_FP64(std::vector<double>{1, 2, 2, 4}, 2);     // singular
    std::vector<double> z = matrix_inversion_no_pivots(std::vector<double>{0, 1, 1, 0}, 2); // zero diagonal, no pivoting
    std::vector<float> f = matrix_inversion_FP32(std::vector<float>{4, 7, 2, 6}, 2);
    std::printf("%d %zu %zu %zu %zu %zu\n", bad, r.size(), q.size(), s.size(), z.size(), f.size());
    if (r.size() == 4) {
        std::printf("%.15f %.15f %.15f %.15f\n", r[0], r[1], r[2], r[3]);
        std::printf("%.3e\n", matrix_multiply(r, a));
    }
    return bad;
}
''')
    exe = tmp_path / "caller64"
    libdir = ROOT / "gpu_matrix_inversion_b200"
    subprocess.run(["g++", "-std=c++14", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-L", str(libdir),
                    "-lmatinv32", f"-Wl,-rpath,{libdir}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                   check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    first = p.stdout.splitlines()[0].split()
    assert first[0] == "0"
    import gpu_matrix_inversion_b200 as m

    if m.device_count() == 0:
        assert first[1:] == ["0", "0", "0", "0", "0"]       # no device -> {} everywhere, never a crash
    else:
        assert first[1:] == ["4", "4", "0", "0", "4"]
        vals = [float(x) for x in p.stdout.splitlines()[1].split()]
        assert np.allclose(vals, [0.6, -0.7, -0.2, 0.4], atol=1e-14)
        assert abs(float(p.stdout.splitlines()[2])) < 1e-12


def test_fp64_python_twins_without_device():
    import gpu_matrix_inversion_b200 as m

    assert m.matrix_inversion_FP64(np.ones(4), 0).size == 0
    assert m.matrix_inversion_FP64(np.ones(5), 3).size == 0
    assert m.matrix_inversion_no_pivots(np.ones(12), 3).size == 0
    assert np.isnan(m.matrix_multiply(np.ones(5), np.ones(5)))
    if m.device_count() == 0:
        assert m.matrix_inversion_FP64(np.eye(2).ravel(), 2).size == 0
        with pytest.raises(m.MatinvError) as e:
            m.invert_f64(np.eye(2))
        assert e.value.code == m.E_NODEVICE
