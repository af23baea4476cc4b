"""GPU tests of the optional 3xTF32 tcgen05 trailing update (MATINV_FLAG_TF32X3, csrc/gj_gemm_tc.cu).

This path is NOT bit-identical to the reference's FMA chain (north_star: "gated by residual"), so the bar here is
  * the kernel alone: exact on operands whose products are exact in TF32, <= 4e-6 of |w| + sum|c u| on random data
    (FP32 SIMT kernel: ~2.5e-7), tiles outside the update untouched;
  * whole inversions: the residual gate passes (<= 1e-5, the tolerance north_star states), the true residual agrees
    with the gate's estimate, the deviation from the FP64 replay stays within a small factor of the FP32 path's;
  * verdicts: singular / non-finite inputs give the FP32 algorithm's verdict (the shim reruns the FP32 schedule);
  * n <= 128 (no trailing update) is bit-identical to the FP32 path.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gj_oracle as o  # noqa: E402


@pytest.fixture(scope="module")
def m():
    import gpu_matrix_inversion_b200 as mod

    assert mod.device_count() >= 1, "CUDA extension loaded but no device: refusing to fall back"
    return mod


@pytest.fixture(scope="module")
def torch():
    import torch as t

    return t


def _skipped(npad, k0, t):
    idx = t.arange(npad, device="cuda") // 128 == k0 // 128
    return idx[:, None] | idx[None, :]


@pytest.mark.parametrize("kk", [0, 1, 3, 4, 7, 8, 15, 16, 17, 64, 127])
def test_onehot_products_are_exact(m, torch, kk):
    """A = e_kk (x) fa, B = e_kk (x) fb with small integers: every product is exact in TF32 and there is one term per sum, so
    the tensor-core result must equal -fa(i) fb(j) exactly -- any mistake in the operand images (core-matrix order,
    descriptor strides, hi/lo placement) shows up here as a wrong element, not as a tolerance question."""
    npad, k0 = 512, 128
    W = torch.zeros((npad, npad), dtype=torch.float32, device="cuda")
    C = torch.zeros((128, npad), dtype=torch.float32, device="cuda")
    U = torch.zeros_like(C)
    fa = (torch.arange(npad, device="cuda") % 251 + 1).float()
    fb = (torch.arange(npad, device="cuda") % 241 + 1).float()
    C[kk] = fa
    U[kk] = fb
    m.debug_trailing_update(W, k0, C, U, mode=1)
    want = -(fa[:, None] * fb[None, :])
    want[_skipped(npad, k0, torch)] = 0.0
    assert torch.equal(W, want)


@pytest.mark.parametrize("npad,k0", [(256, 0), (256, 128), (512, 256), (1152, 0), (1152, 1024), (1152, 384), (2048, 896)])
def test_random_update_against_fp64(m, torch, npad, k0):
    g = torch.Generator(device="cuda").manual_seed(npad * 131 + k0)
    W = (torch.rand((npad, npad), device="cuda", generator=g) - 0.5) * 100.0
    C = torch.rand((128, npad), device="cuda", generator=g) * 2.0 - 1.0
    U = (torch.rand((128, npad), device="cuda", generator=g) - 0.5) * 200.0
    ref = W.double() - C.double().T @ U.double()
    mag = W.double().abs() + C.double().abs().T @ U.double().abs()
    Ws, Wt = W.clone(), W.clone()
    m.debug_trailing_update(Ws, k0, C, U, mode=0)
    m.debug_trailing_update(Wt, k0, C, U, mode=1)
    skip = _skipped(npad, k0, torch)
    assert torch.equal(Wt[skip], W[skip]), "3xTF32 kernel wrote outside the trailing update"
    assert torch.equal(Ws[skip], W[skip])
    live = ~skip
    e_s = float(((Ws.double() - ref).abs() / mag)[live].max())
    e_t = float(((Wt.double() - ref).abs() / mag)[live].max())
    assert e_s < 2e-6, e_s              # FP32 FMA chain (measured ~2.3e-7)
    assert e_t < 4e-6, e_t              # 3xTF32: dropped lo*lo terms + tensor-core accumulation (measured ~6.5e-7)
    assert e_t > 0.0                    # (it is a different arithmetic: equality would mean the wrong kernel ran)


@pytest.mark.parametrize("n,family", [(129, "uniform"), (300, "uniform"), (1000, "diagdom"), (1024, "uniform"),
                                      (1500, "uniform"), (2048, "diagdom"), (2048, "uniform")])
def test_inversion_passes_the_gate(m, torch, n, family):
    A_np = o.uniform(n) if family == "uniform" else o.diagdom(n)
    A = torch.from_numpy(A_np).cuda()
    before = m.tf32x3_status()
    rc, X = m.invert_dev(A, flags=m.FLAG_TF32X3)
    st = m.tf32x3_status()
    assert rc == m.OK
    assert st["inversions"] == before["inversions"] + 1 and not st["fell_back"], st
    assert 0.0 <= st["estimate"] <= m.TF32X3_GATE
    res, _ = m.residual_dev(A, X)
    assert res <= 1e-5                                   # north_star tolerance
    assert 0.1 * res <= st["estimate"] <= 10.0 * res     # the O(n^2) estimate tracks the true residual (4 probes)
    rc32, X32 = m.invert_dev(A)
    res32, _ = m.residual_dev(A, X32)
    assert res <= 16.0 * res32 + 1e-12
    if n <= 1024:   # deviation from the FP64 replay of the same algorithm, next to the FP32 path's
        X64, _, info = o.invert_inplace(A_np.astype(np.float64))
        assert info == 0
        d_t = np.abs(X.cpu().numpy().astype(np.float64) - X64).max()
        d_s = np.abs(X32.cpu().numpy().astype(np.float64) - X64).max()
        assert d_t <= 32.0 * d_s + 1e-12, (d_t, d_s)


def test_small_orders_are_the_fp32_result(m, torch):
    for n in (1, 5, 64, 128):
        A = torch.from_numpy(o.uniform(n)).cuda()
        rc, X = m.invert_dev(A, flags=m.FLAG_TF32X3)
        rc32, X32 = m.invert_dev(A)
        assert rc == rc32 == m.OK
        assert torch.equal(X.view(torch.int32), X32.view(torch.int32))


def test_singular_verdicts_come_from_the_fp32_schedule(m, torch):
    n = 600
    A = o.uniform(n)
    A[7, :] = 0.0
    At = torch.from_numpy(A).cuda()
    rc, _ = m.invert_dev(At, flags=m.FLAG_TF32X3)
    rc32, _ = m.invert_dev(At)
    assert rc == rc32 == m.SINGULAR
    assert m.tf32x3_status()["fell_back"]
    B = o.uniform(n)
    B[0, 0] = np.nan
    rc, _ = m.invert_dev(torch.from_numpy(B).cuda(), flags=m.FLAG_TF32X3)
    assert rc == m.SINGULAR
    assert m.invert(A, flags=m.FLAG_TF32X3) is None      # host entry: None <-> the C++ layer's empty vector


def test_host_entry_and_pivots(m, torch):
    n = 700
    A = o.uniform(n)
    X, piv = m.invert(A, flags=m.FLAG_TF32X3, want_piv=True)
    assert X is not None and not m.tf32x3_status()["fell_back"]
    R = A.astype(np.float64) @ X.astype(np.float64) - np.eye(n)
    assert np.linalg.norm(R) / (n * np.linalg.norm(A) * np.linalg.norm(X)) <= 1e-5
    assert piv.shape == (n,) and np.all(piv >= np.arange(n)) and np.all(piv < n)


def test_aliasing_is_rejected(m, torch):
    A = torch.from_numpy(o.uniform(300)).cuda()
    with pytest.raises(m.MatinvError):
        m.invert_dev(A, A, flags=m.FLAG_TF32X3)
    rc, _ = m.invert_dev(A, A)                            # the FP32 entry still accepts it
    assert rc == m.OK


def test_probe_estimate_tracks_a_bad_inverse(m, torch):
    """The gate has to SEE a bad inverse: perturb a good one and compare the estimate with the full FP64 residual."""
    n = 1024
    A = torch.from_numpy(o.uniform(n)).cuda()
    _, X = m.invert_dev(A)
    g = torch.Generator(device="cuda").manual_seed(5)
    for eps in (1e-6, 1e-3, 1e-1):
        Xp = X * (1.0 + eps * (torch.rand(X.shape, device="cuda", generator=g) - 0.5))
        true, _ = m.residual_dev(A, Xp)
        est = m.probe_residual_dev(A, Xp)
        assert 0.2 * true <= est <= 5.0 * true, (eps, true, est)
    assert m.probe_residual_dev(A, torch.ones_like(X)) > m.TF32X3_GATE   # a nonsense inverse is far above the gate
    Xn = X.clone()
    Xn[3, 3] = float("nan")
    assert not (m.probe_residual_dev(A, Xn) <= m.TF32X3_GATE)     # NaN never passes the gate


def test_full_size_properties_n8192(m, torch):
    """BASELINE-size behaviour through size-independent properties: the gate passes, X*A ~ I on probe vectors, and the
    inverse of the inverse returns to A within the conditioning of the problem."""
    n = 8192
    A = m.generate_dev(n, o.SEED_DIAGDOM + n, "diagdom")
    rc, X = m.invert_dev(A, flags=m.FLAG_TF32X3)
    st = m.tf32x3_status()
    assert rc == m.OK and not st["fell_back"] and st["estimate"] <= 1e-5
    v = torch.ones(n, 4, device="cuda", dtype=torch.float64)
    v[1::2] = -1.0
    r = A.double() @ (X.double() @ v) - v
    assert float(r.norm() / v.norm()) < 1e-3
    rc, A2 = m.invert_dev(X, flags=m.FLAG_TF32X3)
    assert rc == m.OK
    assert float((A2 - A).norm() / A.norm()) < 1e-3


@pytest.mark.parametrize("n", [2048, 4096])
def test_gate_rejects_plausible_but_wrong_inverses(m, torch, n):
    """The acceptance rule must constrain the result at the orders where the tensor-core path matters.  north_star's
    ||AX-I||_F / (n ||A||_F ||X||_F) <= 1e-5 alone does not (an X unrelated to inv(A) scores ~n^-1.5), so the gate also
    bounds the estimate times sqrt(n) (MATINV_TF32X3_GATE_SCALED): the inverse of a DIFFERENT random matrix, a true inverse
    with 1 % relative noise and one with 0.1 % noise are rejected (such noise scores eps / n on the scaled estimate), the FP32 SIMT inverse and the tensor-core inverse of the
    same matrix are accepted."""
    A = m.generate_dev(n, o.SEED_UNIFORM + n, "uniform")
    rc, X = m.invert_dev(A)
    assert rc == m.OK
    ok, est, est_s = m.tf32x3_gate_dev(A, X)
    assert ok and est <= m.TF32X3_GATE and est_s <= m.TF32X3_GATE_SCALED, (est, est_s)
    rc, Xt = m.invert_dev(A, torch.empty_like(A), flags=m.FLAG_TF32X3)
    assert rc == m.OK and not m.tf32x3_status()["fell_back"]
    assert m.tf32x3_gate_dev(A, Xt)[0]
    B = m.generate_dev(n, o.SEED_UNIFORM + n + 12345, "uniform")
    rc, Xb = m.invert_dev(B)
    ok, est, est_s = m.tf32x3_gate_dev(A, Xb)                 # inverse of another matrix: passes 1e-5, fails the scaled bound
    assert not ok and est_s > m.TF32X3_GATE_SCALED, (est, est_s)
    g = torch.Generator(device="cuda").manual_seed(n)
    for rel in (1e-2, 1e-3):
        noise = 1.0 + rel * torch.randn(X.shape, device="cuda", generator=g)
        ok, est, est_s = m.tf32x3_gate_dev(A, X * noise)
        assert not ok, (rel, est, est_s)
