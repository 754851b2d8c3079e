"""b200zk_msm_g1 (host scalars, pinned) at 2^log2n points: wall time per call for each host-chunk setting."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log2n
ctx = zk.Context(0)
lib = zk.load()
srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([12345]), ctx).precompute()
rng = np.random.default_rng(1)
sc = rng.integers(0, 256, size=n * 32, dtype=np.uint8)
sc[31::32] &= 0x1F
h = torch.from_numpy(sc).pin_memory().numpy()
ref = None
for chunks in (1, 2, 3, 4, 6, 8):
    lib.b200zk_msm_set_host_chunks(ctx.handle, chunks)
    r = zk.MultiExp(srs, h)
    ref = ref or r
    assert r == ref
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        zk.MultiExp(srs, h)
        ts.append((time.perf_counter() - t) * 1e3)
    print("chunks=%d  %.2f ms  (%.1f Mpoints/s)" % (chunks, min(ts), n / min(ts) / 1e3))
