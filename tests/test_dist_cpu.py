"""world_size-2 gloo tests (CPU): the host-side logic of the sharded paths — shard layouts, the all-to-all of the
four-step NTT, the partial-sum exchange of the point-range-sharded MSM.  The local compute is replaced by the
oracle (this is a test of the plumbing, not of the CUDA kernels: those are covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noir_backend_using_gnark_b200.dist_ntt import DistributedDomain, ShardLayout
from oracle import bn254 as o
from oracle import cref, dist_ntt_ref


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ntt_worker(rank, world, port, log2n, log2c, full_in, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = ShardLayout(log2n, world, log2c)

        def half_fn(src, dst, half, inverse, decimation, coset):
            out = dist_ntt_ref.half(lay, rank, src.numpy().tobytes(), half, inverse, decimation, coset)
            dst.copy_(torch.from_numpy(np.frombuffer(out, dtype=np.uint8).copy()))

        d = DistributedDomain(1 << log2n, ctx=None, log2c=log2c, half_fn=half_fn)
        out = {}
        for inverse in (0, 1):
            for dec in (o.DIF, o.DIT):
                for coset in (0, 1):
                    x = torch.from_numpy(lay.scatter(full_in, rank, column_block=(dec == o.DIF)).copy())
                    y = (d.FFTInverse if inverse else d.FFT)(x, dec, bool(coset))
                    out[(inverse, dec, coset)] = y.numpy().copy()
        results[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("log2n,log2c", [(6, 3), (7, 4)])
def test_four_step_ntt_layout_world2(log2n, log2c):
    world = 2
    full_in = cref.random_fr(1 << log2n, 0xB2000003 + log2n)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_ntt_worker, args=(world, _free_port(), log2n, log2c, full_in, results), nprocs=world, join=True)
    lay = ShardLayout(log2n, world, log2c)
    for inverse in (0, 1):
        for dec in (o.DIF, o.DIT):
            for coset in (0, 1):
                shards = [results[r][(inverse, dec, coset)] for r in range(world)]
                # DIF output shards are row-block (contiguous chunks), DIT output shards are column-block
                got = lay.gather(shards, column_block=(dec == o.DIT))
                want = cref.ntt(full_in, log2n, inverse, dec, coset, 1)
                assert got.tobytes() == want, (inverse, dec, coset)


def test_shard_layout_roundtrip():
    lay = ShardLayout(10, 4, 6)
    full = cref.random_fr(1 << 10, 3)
    for colb in (True, False):
        shards = [lay.scatter(full, r, colb) for r in range(4)]
        assert lay.gather(shards, colb).tobytes() == full.tobytes()
    # every logical index is owned exactly once
    allidx = np.concatenate([lay.column_block_indices(r) for r in range(4)])
    assert sorted(allidx.tolist()) == list(range(1 << 10))


def _msm_worker(rank, world, port, n, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # point-range shard of one MSM: rank k owns points [k*n, (k+1)*n)
        a, b = 17, 29
        first = o.g1_to_bytes([o.g1_mul(o.G1_GEN, a + rank * n * b)])
        step = o.g1_to_bytes([o.g1_mul(o.G1_GEN, b)])
        pts = cref.g1_arith_progression(first, step, n)
        sc = cref.random_fr(n, 0xB2000001 + rank)
        part = torch.from_numpy(np.frombuffer(cref.msm(pts, sc, n), dtype=np.uint8).copy())
        gathered = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, part)
        total = None
        for g in gathered:
            total = o.g1_add(total, o.g1_from_bytes(g.numpy().tobytes())[0])
        results[rank] = (o.g1_to_bytes([total]), sc.tobytes())
    finally:
        dist.destroy_process_group()


def test_msm_point_range_sharding_world2():
    world, n = 2, 64
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_msm_worker, args=(world, _free_port(), n, results), nprocs=world, join=True)
    scal = []
    for r in range(world):
        scal += o.fr_from_mont_bytes(results[r][1])
    k = sum(s * (17 + 29 * i) for i, s in enumerate(scal)) % o.R_MOD
    want = o.g1_to_bytes([o.g1_mul(o.G1_GEN, k)])
    assert results[0][0] == want and results[1][0] == want
