"""Host logic of the column-sharded inversion on CPU: 2 ranks over gloo with a numpy backend.

What is covered without a GPU: block-cyclic ownership, the per-block message flow (owner factors -> broadcast ->
everybody applies), the deferred column permutation exchanged between ranks, singular status propagation.
The numpy backend below replays the blocked algorithm (oracle A.4) with an emulated FP32 FMA; the 2-rank result
must equal the 1-rank result of the same backend bit for bit, and the oracle within FP32 tolerance.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpu_matrix_inversion_b200.sharded import BLOCK, ShardedInverter, column_gather_list, local_blocks, owner_of
from oracle import gj_oracle as o

f32, f64 = np.float32, np.float64


def fma_neg(a, c, u):
    """fmaf(-c, u, a) emulated through float64 (exact product, one extra rounding that is almost always invisible)."""
    return (a.astype(f64) - c.astype(f64) * u.astype(f64)).astype(f32)


class NumpyShardBackend:
    def __init__(self, n, rank, world, A_full):
        self.n, self.rank, self.world = n, rank, world
        self.blocks = local_blocks(n, rank, world)
        self.W = {J: np.ascontiguousarray(A_full[:, J * BLOCK:(J + 1) * BLOCK]).astype(f32) for J in self.blocks}
        self.piv = np.zeros(n, dtype=np.int32)
        self.info = 0
        self.msg_floats = BLOCK * n + BLOCK + BLOCK + 1

    def new_msg(self):
        return torch.zeros(self.msg_floats, dtype=torch.float32)

    def get_block(self, J):
        return torch.from_numpy(self.W[J].copy())

    def factor(self, J, msg):
        n, k0 = self.n, J * BLOCK
        P = self.W[J]
        kb = P.shape[1]
        C = np.zeros((BLOCK, n), f32); pv = np.zeros(BLOCK, f32); piv = np.zeros(BLOCK, f32); info = 0
        for t in range(kb):
            r = k0 + t
            p = r + int(np.argmax(np.abs(P[r:, t])))
            v = P[p, t]
            piv[t] = p; pv[t] = v
            if v == 0 or not np.isfinite(v):
                info = info or (r + 1)
                v = f32(1.0)
            if p != r:
                P[[r, p]] = P[[p, r]]
                C[:t, [r, p]] = C[:t, [p, r]]
            inv = f32(1.0) / v
            P[r] = P[r] / v
            P[r, t] = inv
            c = P[:, t].copy(); c[r] = 0
            C[t] = c
            upd = fma_neg(P, c[:, None], P[r][None, :])
            upd[:, t] = fma_neg(np.zeros(n, f32), c, np.full(n, inv, f32))
            upd[r] = P[r]
            P[:] = upd
        m = msg.numpy()
        m[:BLOCK * n] = C.ravel(); m[BLOCK * n:BLOCK * n + BLOCK] = pv
        m[BLOCK * n + BLOCK:BLOCK * n + 2 * BLOCK] = piv; m[-1] = info

    def apply(self, J, msg):
        n, k0 = self.n, J * BLOCK
        kb = min(BLOCK, n - k0)
        m = msg.numpy()
        C = m[:BLOCK * n].reshape(BLOCK, n); pv = m[BLOCK * n:BLOCK * n + BLOCK]
        piv = m[BLOCK * n + BLOCK:BLOCK * n + 2 * BLOCK].astype(np.int64)
        self.piv[k0:k0 + kb] = piv[:kb]
        if m[-1] != 0 and self.info == 0:
            self.info = int(m[-1])
        rows = np.arange(n)
        nonpiv = (rows < k0) | (rows >= k0 + kb)
        for Jl, X in self.W.items():
            if Jl == J:
                continue
            for t in range(kb):
                r, p = k0 + t, int(piv[t])
                if p != r:
                    X[[r, p]] = X[[p, r]]
            U = np.zeros((kb, X.shape[1]), f32)
            for t in range(kb):
                r = k0 + t
                u = X[r] / (pv[t] if pv[t] != 0 else f32(1.0))
                U[t] = u; X[r] = u
                for t2 in range(kb):
                    if t2 != t:
                        X[k0 + t2] = fma_neg(X[k0 + t2], np.full_like(u, C[t, k0 + t2]), u)
            acc = X[nonpiv]
            for t in range(kb):
                acc = fma_neg(acc, C[t, nonpiv][:, None], U[t][None, :])
            X[nonpiv] = acc

    def status(self):
        return self.info, self.piv.copy()


def _worker(rank, world, port, n, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    A = {"uniform": o.uniform, "hollow": lambda k: o.hollow(k)[0], "singular": lambda k: np.where(np.arange(k)[:, None] == 7, 0, o.uniform(k)).astype(f32)}[kind](n)
    # the per-block broadcasts may run on their own process group (bench.py: an NCCL group limited to a few CTAs); here a
    # second gloo group over the same ranks, so that the group plumbing of ShardedInverter is exercised on CPU
    bgroup = dist.new_group(ranks=list(range(world))) if world > 1 and kind != "hollow" else None
    inv = ShardedInverter(NumpyShardBackend(n, rank, world, A), dist if world > 1 else None, bcast_group=bgroup)
    info, piv, blocks = inv.invert()
    out = {J: t.numpy().copy() for J, t in blocks.items()} if blocks is not None else None
    q.put((rank, info, piv, out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, n, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    X = np.zeros((n, n), f32)
    info = max(r[1] for r in res)
    for rank, i, piv, out in res:
        if out is not None:
            for J, blk in out.items():
                X[:, J * BLOCK:J * BLOCK + blk.shape[1]] = blk
    return info, res[0][2], X


def test_partition_helpers():
    assert [owner_of(J, 4) for J in range(6)] == [0, 1, 2, 3, 0, 1]
    assert local_blocks(1000, 1, 2) == [1, 3, 5, 7]
    piv = np.array([2, 1, 2, 3], dtype=np.int32)          # swap(0,2) only
    assert list(column_gather_list(piv)) == [2, 1, 0, 3]


@pytest.mark.parametrize("kind", ["uniform", "hollow"])
def test_two_ranks_equal_one_rank_and_oracle(kind):
    n = 300                                                 # 3 column blocks, the last one ragged (44 columns)
    info1, piv1, X1 = _run(1, n, kind)
    info2, piv2, X2 = _run(2, n, kind)
    assert info1 == info2 == 0
    assert np.array_equal(piv1, piv2)
    assert np.array_equal(X1.view(np.uint32), X2.view(np.uint32))          # sharding changes nothing, bit for bit
    A = o.uniform(n) if kind == "uniform" else o.hollow(n)[0]
    Xo, po, io = o.invert_inplace(A)
    assert io == 0 and np.array_equal(piv2, po)
    assert np.allclose(X2, Xo, rtol=1e-4, atol=1e-6 * np.abs(Xo).max())
    res, _ = o.residual(A, X2)
    assert res <= 1e-5


def test_singular_status_reaches_every_rank():
    info, piv, X = _run(2, 300, "singular")
    assert info != 0
