"""CPU tests of the arithmetic model of the optional 3xTF32 trailing update (oracle/tf32x3_model.py): the split is what
cvt.rna.tf32.f32 does, hi + lo recovers 22 bits, and the three-product sum is FP32-grade (2^-21 of sum|c u|) -- the bound
tests/test_gpu_tf32x3.py then holds the tcgen05 kernel to (with room for the tensor core's FP32 accumulation)."""
import numpy as np

from oracle import tf32x3_model as t


def test_rna_tf32_known_values():
    x = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -10, 1.0 + 3 * 2.0 ** -11, -1.0 - 2.0 ** -11, 0.0, -0.0,
                  np.float32(3.14159265), 2.0 ** -126, np.inf, -np.inf], dtype=np.float32)
    r = t.rna_tf32(x)
    # ties go AWAY from zero: 1 + 2^-11 -> 1 + 2^-10, -(1 + 2^-11) -> -(1 + 2^-10)
    assert r[0] == 1.0 and r[1] == np.float32(1.0 + 2.0 ** -10) and r[2] == np.float32(1.0 + 2.0 ** -10)
    assert r[3] == np.float32(1.0 + 2.0 ** -9) and r[4] == np.float32(-1.0 - 2.0 ** -10)
    assert r[5] == 0.0 and np.signbit(r[6]) and r[8] == np.float32(2.0 ** -126)
    assert np.isinf(r[9]) and r[9] > 0 and np.isinf(r[10]) and r[10] < 0
    assert np.all((r.view(np.uint32) & np.uint32(0x1FFF)) == 0)           # 13 low mantissa bits are gone
    assert abs(float(r[7]) - 3.14159265) <= 2.0 ** -10                    # half an ulp of a 10-bit mantissa at 2..4
    assert np.isnan(t.rna_tf32(np.array([np.nan], dtype=np.float32))[0])


def test_split_recovers_22_bits():
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(200000) * 10.0 ** rng.uniform(-6, 6, 200000)).astype(np.float32)
    hi, lo = t.split(x)
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0) and np.all((lo.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    nz = x != 0
    assert np.all(np.abs(x[nz].astype(np.float64) - hi[nz]) <= np.abs(x[nz]) * 2.0 ** -11)             # hi: 11 bits
    rel = np.abs(x[nz].astype(np.float64) - hi[nz].astype(np.float64) - lo[nz]) / np.abs(x[nz])
    assert rel.max() <= 2.0 ** -22                                                                       # hi + lo: 22 bits
    small = np.array([1.0, 3.0, 251.0, 241.0, 2047.0, 1.5 * 2.0 ** 40], dtype=np.float32)                           # <= 11 bits: lo == 0
    assert np.all(t.split(small)[1] == 0)


def test_three_products_are_fp32_grade():
    rng = np.random.default_rng(11)
    K, M, N = 128, 96, 80
    C = rng.uniform(-1, 1, (K, M)).astype(np.float32)
    U = rng.uniform(-100, 100, (K, N)).astype(np.float32)
    exact = C.astype(np.float64).T @ U.astype(np.float64)
    mag = np.abs(C.astype(np.float64)).T @ np.abs(U.astype(np.float64))
    err = np.abs(t.product_model(C, U) - exact) / mag
    assert err.max() <= 2.0 ** -21          # dropped lo*lo (2^-22) + rounding of the two lo parts
    # a single TF32 product (hi*hi only) is ~1000x worse: the variant is NOT "TF32 precision"
    ch, uh = t.split(C)[0].astype(np.float64), t.split(U)[0].astype(np.float64)
    assert (np.abs(ch.T @ uh - exact) / mag).max() > 2.0 ** -13


def test_update_model_matches_fp32_chain_to_rounding():
    """Against the reference's own arithmetic (the FP32 FMA chain of fixColumnKernel applied 128 times) the model differs by
    a few FP32 ulps of the magnitude sum -- the size of the pivot-visible difference the residual gate exists for."""
    rng = np.random.default_rng(3)
    K, M, N = 128, 64, 64
    W = rng.uniform(-50, 50, (M, N)).astype(np.float32)
    C = rng.uniform(-1, 1, (K, M)).astype(np.float32)
    U = rng.uniform(-100, 100, (K, N)).astype(np.float32)
    chain = W.copy()
    for k in range(K):   # w <- fma(-c, u, w), one rounding per step (FP64 product + add rounds like an FMA here)
        chain = (chain.astype(np.float64) - C[k].astype(np.float64)[:, None] * U[k].astype(np.float64)[None, :]).astype(np.float32)
    model = t.trailing_update_model(W, C, U)
    mag = np.abs(W.astype(np.float64)) + np.abs(C.astype(np.float64)).T @ np.abs(U.astype(np.float64))
    assert (np.abs(model.astype(np.float64) - chain) / mag).max() <= 2.0 ** -20
    assert not np.array_equal(model, chain)     # and it is a different arithmetic: bit-equality is not on offer


def test_onehot_patterns_are_exact_in_the_model():
    """The inputs tools/tc_check.cpp and test_gpu_tf32x3.py use to decode layout mistakes: products exact in TF32."""
    n = 256
    fa = (np.arange(n) % 251 + 1).astype(np.float32)
    fb = (np.arange(n) % 241 + 1).astype(np.float32)
    C = np.zeros((128, n), dtype=np.float32)
    U = np.zeros((128, n), dtype=np.float32)
    C[21], U[21] = fa, fb
    assert np.array_equal(t.product_model(C, U), np.outer(fa, fb).astype(np.float64))
