"""Latency of the building blocks at the sizes of the reference's own test programs (tens to thousands of rows)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp

ctx = zk.Context(0)
rng = np.random.default_rng(3)


def images(n):
    a = rng.integers(0, 256, size=n * 32, dtype=np.uint8)
    a[31::32] &= 0x1F
    return a


def wall(fn, reps=20):
    fn()
    ctx.sync()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t) / reps * 1e3


srs = zk.SRS.NewSRS(1 << 14, zkp.fr_to_mont([777]), ctx).precompute()
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
for n in (4, 16, 64, 256, 1024, 4096, 16384):
    sc = torch.from_numpy(images(n)).cuda()
    l0 = ctx.launch_count
    zk.MultiExp(srs, sc, n=n, out=out)
    launches = ctx.launch_count - l0
    ms = wall(lambda: zk.MultiExp(srs, sc, n=n, out=out))
    print("msm n=%6d  %.3f ms  (%d launches)" % (n, ms, launches))
for log2n in (2, 4, 6, 8, 10, 12, 14):
    a = torch.from_numpy(images(1 << log2n)).cuda()
    d = zk.Domain(1 << log2n, ctx)
    ms = wall(lambda: d.FFT(a, zk.DIF, True))
    print("ntt 2^%d coset DIF  %.3f ms" % (log2n, ms))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from prove_bench import synthetic

for log2n in (3, 6, 9, 12):
    c = synthetic(log2n)
    s = zk.SRS.NewSRS((1 << log2n) + 3, zkp.fr_to_mont([777]), ctx).precompute()
    pk = zkp.ProvingKey.SetupRaw(s, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"], c["lro"], ctx)
    bl = images(9)
    l0 = ctx.launch_count
    pk.Prove(c["sol"], bl)
    launches = ctx.launch_count - l0
    ms = wall(lambda: pk.Prove(c["sol"], bl), 10)
    print("prove 2^%d rows  %.3f ms  (%d launches)" % (log2n, ms, launches))
    pk.close()
    s.close()
