"""For each size, time the table-mode MSM for every window width c the precomputation accepts (W = ceil(255/c) windows)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images

ctx = zk.Context(0)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for lg in range(lo, hi + 1):
    n = 1 << lg
    sc = torch.from_numpy(images(n, 7)).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    res = []
    ref = None
    for c in [0] + list(range(max(10, lg - 4), min(23, lg + 5) + 1)):
        srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([12345678901234567890]), ctx)
        srs.precompute(c)
        for _ in range(2):
            zk.MultiExp(srs, sc, n=n, out=out)
        ctx.sync()
        t = time.perf_counter()
        for _ in range(10):
            zk.MultiExp(srs, sc, n=n, out=out)
        ctx.sync()
        ms = (time.perf_counter() - t) / 10 * 1e3
        r = out.cpu().numpy().tobytes()
        ref = ref or r
        assert r == ref
        res.append((c, srs.windows(n), round(ms, 3)))
        srs.close()
    best = min(res[1:], key=lambda x: x[2])
    print("2^%d  auto: W=%d %.3f ms | best: c=%d W=%d %.3f ms | all: %s" % (lg, res[0][1], res[0][2], best[0], best[1], best[2],
                                                                            " ".join("%d:%.2f" % (c, ms) for c, _, ms in res[1:])))
