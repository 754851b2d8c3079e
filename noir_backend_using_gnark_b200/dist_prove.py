"""PLONK prove latency on N GPUs: the prover's commitments (10 MSMs, ~60 % of a 2^22-gate proof) are sharded by point
range over the ranks of one box, everything else stays on rank 0 (BASELINE.json north_star: "MSM shards naturally
by point range across the 8 GPUs of one box, with partial bucket sums combined over NVLink/NCCL").

Rank 0 runs b200zk_plonk_prove with a commitment hook; for every commitment it sends rank k the scalar slice
[k*m, (k+1)*m) over NVLink (NCCL send), all ranks run the single-GPU MSM on their SRS shard (device-generated with
SRS.NewSRS(first=k*m), window table included), the 128-byte partial sums are gathered and summed on rank 0.
Ranks > 0 sit in ShardedCommitter.serve() until rank 0 calls stop()."""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib
from .api import SRS, Context, MultiExp, SumPartials

COMMIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def _view(ptr: int, nbytes: int, device: int):
    """zero-copy torch uint8 view of device memory owned by the library"""
    import torch

    class _V:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

    return torch.as_tensor(_V(), device="cuda:%d" % device)


class ShardedCommitter:
    def __init__(self, ctx: Context, srs_shard: SRS, shard_len: int, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ctx, self.srs, self.m, self.group = ctx, srs_shard, int(shard_len), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = "cuda:%d" % ctx.device
        self.cmd = torch.zeros(1, dtype=torch.int64, device=dev)
        self.part = torch.zeros(128, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(128 * self.world, dtype=torch.uint8, device=dev)
        self.recv = torch.zeros(self.m * 32, dtype=torch.uint8, device=dev) if self.rank else None
        self._cb = COMMIT_FN(self._hook)  # keep the ctypes trampoline alive
        self.error: Optional[BaseException] = None

    def _slice(self, n: int, k: int):
        lo = min(k * self.m, n)
        hi = min((k + 1) * self.m, n)
        return lo, hi

    # ---- rank 0 -----------------------------------------------------------------------------------------
    def attach(self, pk) -> None:
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_plonk_set_commit_hook(pk.handle, self._cb, None))

    def detach(self, pk) -> None:
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_plonk_set_commit_hook(pk.handle, None, None))

    def _hook(self, user, scalars_ptr, n, out_ptr) -> int:
        try:
            torch, dist = self.torch, self.dist
            sc = _view(scalars_ptr, n * 32, self.ctx.device)
            out = _view(out_ptr, 64, self.ctx.device)
            with torch.cuda.stream(self.ctx.torch_stream()):
                self.cmd.fill_(n)
                dist.broadcast(self.cmd, src=0, group=self.group)
                for k in range(1, self.world):
                    lo, hi = self._slice(n, k)
                    if hi > lo:
                        dist.send(sc[lo * 32: hi * 32], dst=k, group=self.group)
                lo, hi = self._slice(n, 0)
                MultiExp(self.srs, sc[lo * 32: hi * 32], n=hi - lo, first_base=0, out=self.part, partial=True)
                dist.all_gather_into_tensor(self.gathered, self.part, group=self.group)
                SumPartials(self.ctx, self.gathered, out=out)
            return 0
        except BaseException as e:  # never let an exception cross the C frame
            self.error = e
            return -2

    def stop(self) -> None:
        with self.torch.cuda.stream(self.ctx.torch_stream()):
            self.cmd.fill_(-1)
            self.dist.broadcast(self.cmd, src=0, group=self.group)
        self.ctx.sync()

    # ---- ranks > 0 --------------------------------------------------------------------------------------
    def serve(self) -> int:
        """worker loop; returns the number of commitments served"""
        torch, dist = self.torch, self.dist
        served = 0
        ext = self.ctx.torch_stream()
        while True:
            with torch.cuda.stream(ext):
                dist.broadcast(self.cmd, src=0, group=self.group)
            self.ctx.sync()
            n = int(self.cmd.item())
            if n < 0:
                return served
            lo, hi = self._slice(n, self.rank)
            with torch.cuda.stream(ext):
                if hi > lo:
                    dist.recv(self.recv[: (hi - lo) * 32], src=0, group=self.group)
                MultiExp(self.srs, self.recv, n=hi - lo, first_base=0, out=self.part, partial=True)
                dist.all_gather_into_tensor(self.gathered, self.part, group=self.group)
            served += 1
