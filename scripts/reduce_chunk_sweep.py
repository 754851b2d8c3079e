"""MSM time per reduction chunk size (b200zk_msm_set_reduce_chunk) at a few sizes, window table."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images

ctx = zk.Context(0)
lib = zk.load()
for lg in [int(a) for a in sys.argv[1:]] or [18, 19, 20, 22, 24]:
    n = 1 << lg
    srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([424242]), ctx).precompute()
    sc = torch.from_numpy(images(n, 9)).cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    res = []
    for cl in (3, 4, 5):
        lib.b200zk_msm_set_reduce_chunk(ctx.handle, cl)
        for _ in range(2):
            zk.MultiExp(srs, sc, n=n, out=out)
        ctx.sync()
        t = time.perf_counter()
        for _ in range(8):
            zk.MultiExp(srs, sc, n=n, out=out)
        ctx.sync()
        res.append("2^%d chunk: %.3f ms" % (cl, (time.perf_counter() - t) / 8 * 1e3))
    lib.b200zk_msm_set_reduce_chunk(ctx.handle, 0)
    print("2^%d points:" % lg, " | ".join(res))
    srs.close()
