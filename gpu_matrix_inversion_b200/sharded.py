"""Column-sharded inversion of one large matrix across the GPUs of a node -- host side.

One process per GPU (torchrun).  The n x n matrix is dealt by column blocks of 128: block J lives on rank
J % world.  Per block step the owner factors its panel (`matinv_shard_factor`, which writes the broadcast message
in place), the message -- 128 x n multipliers, 128 pivot rows/values, the panel's net row permutation -- is
broadcast with torch.distributed (NCCL over NVLink), and every rank applies it to its own columns
(`matinv_shard_apply`: row interchanges, row-block recurrence, trailing GEMM).  After the last block the deferred
column permutation moves columns between ranks with one all-to-all.

The reference is single-device (/root/reference/Matlab/mat_inv_32/mat_inv_32/mat_inv_32.cpp:239-250); this is the
multi-GPU row of SURVEY.md s.8(e).  Because every rank runs the same kernels on its columns, the sharded result is
bit-identical to the single-GPU result.

The math lives behind a small backend interface so that the host logic (ownership, message flow, column exchange)
is testable on CPU with gloo and a numpy backend (tests/test_sharded_host.py); `CudaShardBackend` is the product.
"""
from __future__ import annotations

import ctypes

import numpy as np

BLOCK = 128


def owner_of(J: int, world: int) -> int:
    return J % world


def local_blocks(n: int, rank: int, world: int):
    nblk = (n + BLOCK - 1) // BLOCK
    return [J for J in range(nblk) if J % world == rank]


def column_gather_list(piv: np.ndarray) -> np.ndarray:
    """colsrc with X[:, j] = M[:, colsrc[j]]: net effect of `for r = n-1..0: swap columns r, piv[r]`
    (device twin: csrc/gj_finish.cu:colperm_kernel)."""
    n = len(piv)
    idx = np.arange(n, dtype=np.int64)
    for r in range(n - 1, -1, -1):
        p = int(piv[r])
        if p != r:
            idx[r], idx[p] = idx[p], idx[r]
    return idx


class CudaShardBackend:
    """Thin wrapper of the matinv_shard_* C-ABI for one rank."""

    def __init__(self, n: int, rank: int, world: int, device):
        import torch

        import gpu_matrix_inversion_b200 as m

        self.m, self.torch, self.n, self.rank, self.world, self.device = m, torch, n, rank, world, device
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            m._check(m.lib.matinv_shard_create(n, rank, world, ctypes.byref(h)))
        self.h = h
        self.msg_bytes = int(m.lib.matinv_shard_panel_bytes(n))
        self.blocks = local_blocks(n, rank, world)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def new_msg(self):
        return self.torch.zeros(self.msg_bytes, dtype=self.torch.uint8, device=self.device)

    def generate(self, seed: int, kind: str = "uniform"):
        self.m._check(self.m.lib.matinv_shard_generate(self.h, seed, 1 if kind == "diagdom" else 0, self._stream()))

    def set_block(self, J: int, t):
        """t: (n, <=128) float32 tensor (any device), row-major."""
        t = t.contiguous()
        self.m._check(self.m.lib.matinv_shard_set_block(self.h, J, ctypes.c_void_p(t.data_ptr()), t.stride(0), self._stream()))
        self.torch.cuda.current_stream(self.device).synchronize()

    def get_block(self, J: int):
        ncols = min(BLOCK, self.n - J * BLOCK)
        t = self.torch.empty((self.n, ncols), dtype=self.torch.float32, device=self.device)
        self.m._check(self.m.lib.matinv_shard_get_block(self.h, J, ctypes.c_void_p(t.data_ptr()), t.stride(0), self._stream()))
        return t

    def factor(self, J: int, msg):
        self.m._check(self.m.lib.matinv_shard_factor(self.h, J, ctypes.c_void_p(msg.data_ptr()), self._stream()))

    def apply(self, J: int, msg):
        self.m._check(self.m.lib.matinv_shard_apply(self.h, J, ctypes.c_void_p(msg.data_ptr()), self._stream()))

    def apply_only(self, J: int, msg, block: int):
        """Update of panel J restricted to local block `block` (look-ahead: the next panel's columns first)."""
        self.m._check(self.m.lib.matinv_shard_apply_ex(self.h, J, ctypes.c_void_p(msg.data_ptr()), self._stream(), 1, block))

    def apply_except(self, J: int, msg, block: int):
        self.m._check(self.m.lib.matinv_shard_apply_ex(self.h, J, ctypes.c_void_p(msg.data_ptr()), self._stream(), 2, block))

    lookahead_capable = True

    def status(self):
        info = ctypes.c_int(0)
        piv = np.empty(self.n, dtype=np.int32)
        rc = self.m.lib.matinv_shard_status(self.h, ctypes.byref(info), piv.ctypes.data, self._stream())
        self.m._check(rc)
        return info.value, piv

    def close(self):
        if self.h:
            self.m.lib.matinv_shard_destroy(self.h)
            self.h = None


class ShardedInverter:
    """Host-side schedule of the column-sharded inversion (one instance per rank)."""

    def __init__(self, backend, dist=None, bcast_group=None):
        self.b = backend
        self.dist = dist  # torch.distributed or None when world == 1
        # process group of the per-block broadcasts (None = default group).  bench.py passes an NCCL group limited to two
        # CTAs: a receiver's broadcast kernel spins for a whole block step, and every CTA it holds is an SM slot the trailing
        # GEMM does not get (csrc/gj_multi.cu has the measurement)
        self.bcast_group = bcast_group
        self.n, self.rank, self.world = backend.n, backend.rank, backend.world
        self.nblk = (self.n + BLOCK - 1) // BLOCK
        self.msg = [backend.new_msg(), backend.new_msg()]

    def factorize(self, lookahead=None):
        """All block steps.  Returns (info, piv); the shards then hold M = inv(P A) column-wise."""
        if lookahead is None:
            import os

            # measured on one B200 (world = 1): 11.99 s -> 11.11 s at n = 65536, 1636 -> 1558 ms at n = 32768
            lookahead = getattr(self.b, "lookahead_capable", False) and os.environ.get("MATINV_SHARD_LOOKAHEAD", "1") != "0"
        if lookahead:
            return self._factorize_lookahead()
        for J in range(self.nblk):
            own = owner_of(J, self.world)
            msg = self.msg[J & 1]
            if own == self.rank:
                self.b.factor(J, msg)
            if self.dist is not None and self.world > 1:
                self.dist.broadcast(msg, src=own, group=self.bcast_group)
            self.b.apply(J, msg)
        return self.b.status()

    def _factorize_lookahead(self):
        """Same work, re-ordered: the owner of panel J+1 updates that panel's columns first, factors it on a
        high-priority side stream and starts its broadcast while everybody (itself included) is still applying
        panel J to the remaining columns; the receivers post the broadcast early on their side stream."""
        import torch

        import os

        la_mode = os.environ.get("MATINV_SHARD_LA_MODE", "prio")    # tuning aid: prio | noprio | serial
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = torch.cuda.Stream(priority=-1 if la_mode == "prio" else 0)
        side = main if la_mode == "serial" else self._side
        multi = self.dist is not None and self.world > 1
        msg0 = self.msg[0]
        if owner_of(0, self.world) == self.rank:
            self.b.factor(0, msg0)
        if multi:
            self.dist.broadcast(msg0, src=owner_of(0, self.world), group=self.bcast_group)
        for J in range(self.nblk):
            msg = self.msg[J & 1]
            nxt = J + 1
            if nxt >= self.nblk:
                self.b.apply(J, msg)
                break
            nmsg = self.msg[nxt & 1]
            own_n = owner_of(nxt, self.world)
            top = torch.cuda.Event()
            top.record(main)                      # everything that read nmsg (apply of panel J-1) is before this point
            work = None
            if own_n == self.rank:
                self.b.apply_only(J, msg, nxt)
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ready)
                    self.b.factor(nxt, nmsg)
                    if multi:
                        work = self.dist.broadcast(nmsg, src=own_n, async_op=True, group=self.bcast_group)
                self.b.apply_except(J, msg, nxt)
            else:
                if multi:
                    with torch.cuda.stream(side):
                        side.wait_event(top)
                        work = self.dist.broadcast(nmsg, src=own_n, async_op=True, group=self.bcast_group)
                self.b.apply(J, msg)
            if work is not None:
                work.wait()                       # main stream waits for the broadcast (no host block)
            done = torch.cuda.Event()
            done.record(side)
            main.wait_event(done)
        return self.b.status()

    def exchange_columns(self, piv):
        """Deferred column permutation across shards: returns {J: (n, ncols) tensor} of X for the local blocks."""
        import torch

        n, world, rank = self.n, self.world, self.rank
        colsrc = column_gather_list(piv)
        mine = self.b.blocks
        blocks = [self.b.get_block(J) for J in mine]      # columns of M held here
        dev = blocks[0].device if blocks else self.msg[0].device
        Lm = torch.cat(blocks, dim=1) if blocks else torch.empty((n, 0), dtype=torch.float32, device=dev)

        j = np.arange(n, dtype=np.int64)
        src_rank = (colsrc // BLOCK) % world
        dst_rank = (j // BLOCK) % world
        local_index = lambda c: (c // BLOCK // world) * BLOCK + c % BLOCK  # noqa: E731  (position inside Lm / Om)

        send, recv, recv_slots = [], [], []
        for d in range(world):
            sel = colsrc[(src_rank == rank) & (dst_rank == d)]             # in increasing destination column order
            idx = torch.from_numpy(local_index(sel)).to(dev)
            send.append(Lm.index_select(1, idx).t().contiguous())
            slots = j[(dst_rank == rank) & (src_rank == d)]
            recv_slots.append(torch.from_numpy(local_index(slots)).to(dev))
            recv.append(torch.empty((len(slots), n), dtype=torch.float32, device=dev))
        if self.dist is not None and world > 1:
            if self.dist.get_backend() == "nccl":
                self.dist.all_to_all(recv, send)
            else:  # gloo (CPU tests) has no all_to_all: pairwise exchange
                recv[rank].copy_(send[rank])
                reqs = []
                for peer in range(world):
                    if peer != rank:
                        reqs.append(self.dist.isend(send[peer], dst=peer))
                        reqs.append(self.dist.irecv(recv[peer], src=peer))
                for q in reqs:
                    q.wait()
        else:
            recv = send
        Om = torch.empty_like(Lm)
        for s_ in range(world):
            if recv[s_].shape[0]:
                Om.index_copy_(1, recv_slots[s_], recv[s_].t())
        out, off = {}, 0
        for J in mine:
            w = min(BLOCK, n - J * BLOCK)
            out[J] = Om[:, off:off + w]
            off += w
        return out

    def invert(self):
        info, piv = self.factorize()
        if info != 0:
            return info, piv, None
        return info, piv, self.exchange_columns(piv)
