"""Column-sharded path on the GPU: the ranks of one node emulated in ONE process on cuda:0 (sequential launches, no
kernel waits on another), checked bit for bit against the single-GPU result and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gj_oracle as o  # noqa: E402


def bits(x):
    return np.ascontiguousarray(x).view(np.uint32)


def run_emulated(n, world, A, split=False):
    import torch

    from gpu_matrix_inversion_b200.sharded import BLOCK, CudaShardBackend, ShardedInverter, column_gather_list

    dev = torch.device("cuda", 0)
    backs = [CudaShardBackend(n, r, world, dev) for r in range(world)]
    At = torch.from_numpy(A).to(dev)
    for b in backs:
        for J in b.blocks:
            b.set_block(J, At[:, J * BLOCK:(J + 1) * BLOCK])
    msg = backs[0].new_msg()
    nblk = (n + BLOCK - 1) // BLOCK
    for J in range(nblk):
        backs[J % world].factor(J, msg)
        for b in backs:
            if split and J + 1 < nblk and (J + 1) % world == b.rank:   # the look-ahead split of the same update
                b.apply_only(J, msg, J + 1)
                b.apply_except(J, msg, J + 1)
            else:
                b.apply(J, msg)
    infos = [b.status() for b in backs]
    info, piv = infos[0]
    assert all(i[0] == info and np.array_equal(i[1], piv) for i in infos)
    M = torch.empty((n, n), dtype=torch.float32, device=dev)
    for b in backs:
        for J in b.blocks:
            blk = b.get_block(J)
            M[:, J * BLOCK:J * BLOCK + blk.shape[1]] = blk
    X = M[:, torch.from_numpy(column_gather_list(piv)).to(dev)].cpu().numpy() if info == 0 else None
    for b in backs:
        b.close()
    return info, piv, X


@pytest.mark.parametrize("n,world", [(300, 1), (300, 2), (1000, 3), (1024, 4), (2048, 8)])
def test_sharded_equals_single_gpu_bitwise(n, world):
    import gpu_matrix_inversion_b200 as m

    A = o.uniform(n)
    info, piv, X = run_emulated(n, world, A)
    Xs, pivs = m.invert(A, want_piv=True)
    assert info == 0 and Xs is not None
    assert np.array_equal(piv, pivs)
    assert np.array_equal(bits(X), bits(Xs))
    Xo, po, io = o.invert_inplace(A)
    assert np.array_equal(piv, po) and np.array_equal(bits(X), bits(Xo))


def test_sharded_singular_and_generator():
    import torch

    from gpu_matrix_inversion_b200.sharded import BLOCK, CudaShardBackend

    n = 640
    A = o.uniform(n); A[7] = 0.0
    info, piv, X = run_emulated(n, 2, A)
    assert info != 0 and X is None
    # the shard generator fills the same bits as the unsharded one
    b = CudaShardBackend(n, 1, 2, torch.device("cuda", 0))
    b.generate(o.SEED_UNIFORM + n, "uniform")
    G = o.uniform(n)
    for J in b.blocks:
        assert np.array_equal(bits(b.get_block(J).cpu().numpy()), bits(G[:, J * BLOCK:(J + 1) * BLOCK]))
    b.close()


@pytest.mark.parametrize("n,world", [(1000, 1), (1000, 2), (1536, 3)])
def test_lookahead_split_is_bitwise_identical(n, world):
    A = o.uniform(n)
    i0, p0, X0 = run_emulated(n, world, A, split=False)
    i1, p1, X1 = run_emulated(n, world, A, split=True)
    assert i0 == i1 == 0 and np.array_equal(p0, p1)
    assert np.array_equal(bits(X0), bits(X1))


def test_single_rank_sharded_inverter_with_lookahead_streams():
    """ShardedInverter's look-ahead schedule (side stream, events) on one rank against the single-GPU result."""
    import torch

    import gpu_matrix_inversion_b200 as m
    from gpu_matrix_inversion_b200.sharded import BLOCK, CudaShardBackend, ShardedInverter

    n = 1280
    A = o.uniform(n)
    dev = torch.device("cuda", 0)
    b = CudaShardBackend(n, 0, 1, dev)
    At = torch.from_numpy(A).to(dev)
    for J in b.blocks:
        b.set_block(J, At[:, J * BLOCK:(J + 1) * BLOCK])
    inv = ShardedInverter(b, None)
    info, piv, blocks = inv.invert()
    torch.cuda.synchronize()
    assert info == 0
    X = torch.cat([blocks[J] for J in b.blocks], dim=1).cpu().numpy()
    Xs, pivs = m.invert(A, want_piv=True)
    assert np.array_equal(piv, pivs) and np.array_equal(bits(X), bits(Xs))
    b.close()
