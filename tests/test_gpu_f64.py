"""GPU parity of the FP64 entry points against oracle/gj_oracle.c:gj_inplace_f64,
through the C-ABI.  Bar: bit-exact inverse and pivot sequence (integer view of the doubles); residual
||AX - I||_F / (n ||A|| ||X||) <= 1e-13 at sizes the oracle does not reach in seconds."""
import numpy as np
import pytest

from oracle import gj_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import gpu_matrix_inversion_b200 as mod

    if mod.device_count() == 0:
        pytest.skip("no CUDA device")
    return mod


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def uniform64(n, seed):
    return np.random.default_rng(seed).random((n, n)) * 100.0


@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 100, 257, 500, 1024])
def test_f64_bit_exact_with_pivoting(m, n):
    A = uniform64(n, 6400 + n)
    X, piv = m.invert_f64(A, want_piv=True)
    Xo, po, io = o.invert_inplace(A)
    assert io == 0 and X is not None
    assert np.array_equal(piv, po)
    assert np.array_equal(bits(X), bits(Xo))


def test_f64_upcast_of_the_fp32_workload(m):
    """The FP32 bench generator, widened: same pivots as the FP64 oracle (not necessarily as the FP32 run)."""
    n = 384
    A = o.generate(n, o.SEED_UNIFORM + n, "uniform").astype(np.float64)
    X, piv = m.invert_f64(A, want_piv=True)
    Xo, po, io = o.invert_inplace(A)
    assert io == 0 and np.array_equal(piv, po) and np.array_equal(bits(X), bits(Xo))


@pytest.mark.parametrize("n", [1, 2, 65, 300])
def test_f64_no_pivot_bit_exact(m, n):
    A = o.generate(n, o.SEED_DIAGDOM + n, "diagdom").astype(np.float64)
    X, piv = m.invert_f64(A, nopivot=True, want_piv=True)
    Xo, po, io = o.invert_inplace(A, flags=o.NOPIVOT)
    assert io == 0 and X is not None
    assert np.array_equal(piv, np.arange(n)) and np.array_equal(po, np.arange(n))
    assert np.array_equal(bits(X), bits(Xo))


def test_f64_singular_and_non_finite_inputs(m):
    n = 96
    A = uniform64(n, 1)
    Z = A.copy(); Z[7] = 0.0
    assert m.invert_f64(Z) is None and o.invert_inplace(Z)[2] != 0
    assert m.invert_f64(np.zeros((n, n))) is None
    N = A.copy(); N[0, 0] = np.nan
    assert m.invert_f64(N) is None and o.invert_inplace(N)[2] != 0
    I = A.copy(); I[5, 9] = np.inf
    assert (m.invert_f64(I) is None) == (o.invert_inplace(I)[2] != 0)
    # a permutation matrix needs its row interchanges: fine with pivoting, singular without
    P = np.eye(n)[::-1].copy()
    Xp = m.invert_f64(P)
    assert Xp is not None and np.array_equal(Xp, P.T)
    assert m.invert_f64(P, nopivot=True) is None
    # ties in magnitude: lowest row wins, sign kept
    T = np.random.default_rng(3).integers(-2, 3, size=(n, n)).astype(np.float64)
    X, piv = m.invert_f64(T, want_piv=True)
    Xo, po, io = o.invert_inplace(T)
    assert (X is None) == (io != 0)
    if io == 0:
        assert np.array_equal(piv, po) and np.array_equal(bits(X), bits(Xo))


def test_f64_device_entry_vector_twins_and_defect(m):
    import torch

    n = 200
    A = uniform64(n, 77)
    d = torch.from_numpy(A).cuda()
    piv = torch.empty(n, dtype=torch.int32, device="cuda")
    rc, X = m.invert_f64_dev(d, piv=piv)
    assert rc == m.OK
    Xo, po, _ = o.invert_inplace(A)
    assert np.array_equal(piv.cpu().numpy(), po) and np.array_equal(bits(X.cpu().numpy()), bits(Xo))
    rc, X2 = m.invert_f64_dev(d.clone(), X=None)          # fresh output tensor
    assert rc == m.OK and torch.equal(X, X2)
    v = m.matrix_inversion_FP64(A.ravel(), n)
    assert v.size == n * n and np.array_equal(bits(v.reshape(n, n)), bits(Xo))
    assert m.matrix_inversion_FP64(np.r_[A.ravel(), [1.0, 2.0]], n).size == n * n     # size = N*N + k, k < N: tail ignored
    Zs = A.copy(); Zs[3] = 0.0
    assert m.matrix_inversion_FP64(Zs.ravel(), n).size == 0
    # the reference's verification step: sqrt(n) - ||A X||_F (matrix_multiply.cpp:194-200), first argument = right factor
    e = m.matrix_multiply(Xo.ravel(), A.ravel())
    assert abs(e) < 1e-9
    host = np.sqrt(n) - np.linalg.norm(A @ Xo)
    assert abs(e - host) < 1e-10


def test_f64_residual_at_n2048(m):
    import torch

    n = 2048
    A = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 100
    rc, X = m.invert_f64_dev(A)
    assert rc == m.OK
    R = A @ X - torch.eye(n, dtype=torch.float64, device="cuda")
    rel = float(R.norm() / (n * A.norm() * X.norm()))
    assert rel <= 1e-13
    rc, A2 = m.invert_f64_dev(X)
    assert rc == m.OK and float((A2 - A).abs().max() / A.abs().max()) <= 1e-7


@pytest.mark.parametrize("n", [63, 64, 65, 130, 777, 1500])
def test_f64_blocked_equals_unblocked(m, n):
    """The default (blocked, 64-column panels) and the unblocked schedule are the same arithmetic in a different order of
    independent operations: same bytes, with and without pivoting."""
    A = uniform64(n, 9100 + n)
    Xb, pb = m.invert_f64(A, want_piv=True)
    Xu, pu = m.invert_f64(A, want_piv=True, flags=m.FLAG_UNBLOCKED)
    assert Xb is not None and np.array_equal(pb, pu) and np.array_equal(bits(Xb), bits(Xu))
    D = A + np.eye(n) * 100.0 * n
    Db = m.invert_f64(D, nopivot=True)
    Du = m.invert_f64(D, nopivot=True, flags=m.FLAG_UNBLOCKED)
    assert np.array_equal(bits(Db), bits(Du))
    if n <= 777:
        Xo, po, io = o.invert_inplace(A)
        assert io == 0 and np.array_equal(pu, po) and np.array_equal(bits(Xu), bits(Xo))
