"""Multi-GPU fr NTT (N >= 2^24): four-step decomposition with one all-to-all, one process per GPU.

Replaces (*fft.Domain).FFT / FFTInverse of gnark-crypto v0.9.1 for transforms that are sharded over the GPUs of one
box (BASELINE.json north_star: "NTTs of 2^24 and above use a four-step decomposition with an NVLink all-to-all").

The logical N-point vector is an R x C matrix (index = r*C + c).  Because the ordinary radix-2 DIF/DIT stage
twiddles w^(j*2^s) already contain the inter-block factors, no separate twiddle pass is needed: the stages with
stride >= C only couple elements of one column, the stages with stride < C only couple elements of one row, so

  DIF (natural in -> bit-reversed out):  column-block shards --[strides >= C, local]--> all-to-all (transpose)
                                         --[strides < C, local]--> row-block shards = contiguous chunks of the output
  DIT (bit-reversed in -> natural out):  row-block shards (contiguous chunks of the input) --[strides < C]-->
                                         all-to-all --[strides >= C]--> column-block shards of the natural output

Shard layouts are documented at b200zk_ntt_dist_half_dev in include/b200zk.h.  The local halves run in
libb200zk.so; this module only owns the layout arithmetic and the collective (torch.distributed all_to_all_single:
NCCL over NVLink on GPUs; gloo in the CPU tests, which inject a checker for the local halves).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import DIF, DIT


def _log2(x: int) -> int:
    n = x.bit_length() - 1
    assert 1 << n == x, "power of two required"
    return n


class ShardLayout:
    """Index arithmetic of the shard layouts (pure Python / numpy; shared by product and tests)."""

    def __init__(self, log2n: int, world: int, log2c: Optional[int] = None):
        self.log2n = log2n
        self.world = world
        self.log2g = _log2(world)
        if log2c is None:
            log2c = max(log2n - 8, self.log2g + 2, (log2n + 1) // 2)
        self.log2c = log2c
        assert log2c >= self.log2g + 2 and log2n - log2c >= self.log2g, "transform too small for this many ranks"
        self.R = 1 << (log2n - log2c)
        self.C = 1 << log2c
        self.C_loc = self.C >> self.log2g
        self.R_loc = self.R >> self.log2g
        self.local = 1 << (log2n - self.log2g)

    def column_block_indices(self, rank: int) -> np.ndarray:
        """logical index of every element of rank's column-block shard X[r][c_lo]"""
        r = np.arange(self.R, dtype=np.int64)[:, None]
        c = rank * self.C_loc + np.arange(self.C_loc, dtype=np.int64)[None, :]
        return (r * self.C + c).reshape(-1)

    def row_block_indices(self, rank: int) -> np.ndarray:
        """logical index of every element of rank's row-block shard Y[r_lo][c] (a contiguous chunk)"""
        return rank * self.local + np.arange(self.local, dtype=np.int64)

    def scatter(self, full: np.ndarray, rank: int, column_block: bool) -> np.ndarray:
        """shard of a full vector given as a (N, 32) uint8 array"""
        idx = self.column_block_indices(rank) if column_block else self.row_block_indices(rank)
        return np.ascontiguousarray(full.reshape(-1, 32)[idx]).reshape(-1)

    def gather(self, shards: list, column_block: bool) -> np.ndarray:
        out = np.zeros((1 << self.log2n, 32), dtype=np.uint8)
        for rank, sh in enumerate(shards):
            idx = self.column_block_indices(rank) if column_block else self.row_block_indices(rank)
            out[idx] = np.asarray(sh).reshape(-1, 32)
        return out.reshape(-1)


class DistributedDomain:
    """fft.NewDomain(2^log2n) sharded over the ranks of a torch.distributed process group."""

    def __init__(self, m: int, ctx=None, group=None, log2c: Optional[int] = None,
                 half_fn: Optional[Callable] = None, p2p: bool = False):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        log2n = 0
        while (1 << log2n) < m:
            log2n += 1
        self.layout = ShardLayout(log2n, self.world, log2c)
        self.Cardinality = 1 << log2n
        self.ctx = ctx
        self._half_fn = half_fn or self._device_half
        self._scratch = None
        self.p2p = bool(p2p)
        self._peer_ptrs = None
        if self.p2p:
            self._setup_p2p()

    # -- fused exchange: peer-mapped exchange buffers (CUDA IPC), one per rank ------------------------
    def _setup_p2p(self) -> None:
        import ctypes as C

        import torch

        assert self.world <= 8, "the fused exchange addresses at most 8 peers"
        lib, h = self.ctx.lib, self.ctx.handle
        nbytes = self.layout.local * 32
        own = C.c_void_p()
        _lib.check(h, lib.b200zk_dev_alloc(h, nbytes, C.byref(own)))
        self._own_buf = own
        handle = (C.c_ubyte * 64)()
        _lib.check(h, lib.b200zk_ipc_export(h, own, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda:%d" % self.ctx.device)
        allh = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allh, mine, group=self.group)
        ptrs = (C.c_void_p * self.world)()
        self._imported = []
        for k in range(self.world):
            if k == self.rank:
                ptrs[k] = own.value
            else:
                raw = (C.c_ubyte * 64)(*allh[k].cpu().tolist())
                out = C.c_void_p()
                _lib.check(h, lib.b200zk_ipc_import(h, raw, C.byref(out)))
                ptrs[k] = out.value
                self._imported.append(out)
        self._peer_ptrs = ptrs

        class _View:  # zero-copy torch view of the library-owned exchange buffer
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (own.value, False), "version": 2}

        self._exchange_buf = torch.as_tensor(_View(), device="cuda:%d" % self.ctx.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda:%d" % self.ctx.device)

    def _stream_barrier(self) -> None:
        """cross-rank barrier in stream order (a 4-byte all-reduce on the library's stream)"""
        import torch

        with torch.cuda.stream(self.ctx.torch_stream()):
            self.dist.all_reduce(self._flag, group=self.group)

    def _run_p2p(self, x, inverse: int, decimation: int, coset: int):
        lay = self.layout
        lib, h = self.ctx.lib, self.ctx.handle
        self._stream_barrier()   # every rank is done with its exchange buffer from the previous transform
        rc = lib.b200zk_ntt_dist_half0_p2p_dev(h, x.data_ptr(), self._peer_ptrs, lay.log2n, lay.log2g, self.rank,
                                               lay.log2c, inverse, decimation, coset)
        _lib.check(h, rc)
        self._stream_barrier()   # all peers' stores have landed
        ex = self._exchange_buf
        if decimation == DIF:
            self._device_half(ex, x, 1, inverse, decimation, coset)   # reads Z from the exchange buffer, writes Y
            return x
        self._device_half(ex, ex, 1, inverse, decimation, coset)      # in place on X
        return ex

    def close(self) -> None:
        if self._peer_ptrs is not None:
            lib, h = self.ctx.lib, self.ctx.handle
            self.ctx.sync()
            self.dist.barrier(group=self.group)
            for p in self._imported:
                lib.b200zk_ipc_close(h, p)
            self.dist.barrier(group=self.group)
            lib.b200zk_dev_free(h, self._own_buf)
            self._peer_ptrs = None

    # -- local halves (libb200zk.so) ---------------------------------------------------------------
    def _device_half(self, src, dst, half: int, inverse: int, decimation: int, coset: int) -> None:
        lay = self.layout
        lib, h = self.ctx.lib, self.ctx.handle
        rc = lib.b200zk_ntt_dist_half_dev(h, src.data_ptr(), dst.data_ptr(), lay.log2n, lay.log2g, self.rank,
                                          lay.log2c, half, inverse, decimation, coset)
        _lib.check(h, rc)

    def _buffer_like(self, x):
        import torch

        if self._scratch is None or self._scratch.shape != x.shape or self._scratch.device != x.device:
            self._scratch = torch.empty_like(x)
        return self._scratch

    def _exchange(self, recv, send) -> None:
        """the four-step transpose: chunk j of `send` goes to rank j, chunk k of `recv` comes from rank k"""
        if send.is_cuda:
            import torch

            # order the collective after the library's stream and the next half after the collective
            ext = self.ctx.torch_stream()
            with torch.cuda.stream(ext):
                self.dist.all_to_all_single(recv, send, group=self.group)
        else:
            self.dist.all_to_all_single(recv, send, group=self.group)

    def _run(self, x, inverse: int, decimation: int, coset: bool):
        assert x.numel() * x.element_size() == self.layout.local * 32
        coset = int(bool(coset))
        if self.p2p:
            return self._run_p2p(x, inverse, decimation, coset)
        tmp = self._buffer_like(x)
        if decimation == DIF:
            self._half_fn(x, x, 0, inverse, decimation, coset)       # strides >= C, column-block, in place
            self._exchange(tmp, x)                                    # tmp = Z[peer][r_lo][c_lo]
            self._half_fn(tmp, x, 1, inverse, decimation, coset)     # strides < C: reads Z, writes Y into x
            return x
        self._half_fn(x, tmp, 0, inverse, decimation, coset)         # strides < C on Y, writes Z into tmp
        self._exchange(x, tmp)                                        # x = X[r][c_lo]
        self._half_fn(x, x, 1, inverse, decimation, coset)           # strides >= C, in place
        return x

    def FFT(self, x, decimation: int, coset: bool = False):
        """Sharded domain.FFT.  x: this rank's shard (column-block for DIF, row-block for DIT); returns the tensor
        holding the result shard (row-block for DIF, column-block for DIT) — x itself, or with p2p=True and DIT the
        domain's exchange buffer (valid until the next transform).  x is clobbered either way."""
        return self._run(x, 0, decimation, coset)

    def FFTInverse(self, x, decimation: int, coset: bool = False):
        return self._run(x, 1, decimation, coset)
