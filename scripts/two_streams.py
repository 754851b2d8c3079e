"""Do independent MSMs overlap usefully when issued from two contexts (two streams, two workspaces) of one GPU?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
a, b = zk.Context(0), zk.Context(0)
lib = zk.load()
srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([987654321]), a).precompute()
sc = [torch.from_numpy(images(n, 7 + i)).cuda() for i in range(2)]
outs = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(2)]
torch.cuda.synchronize()


def run(ctxs, reps):
    for i in range(reps):
        c = ctxs[i % len(ctxs)]
        rc = lib.b200zk_msm_g1_dev(c.handle, srs.handle, 0, sc[i % 2].data_ptr(), n, outs[i % 2].data_ptr(), 0)
        assert rc == 0
    for c in ctxs:
        c.sync()


for ctxs, name in (([a], "one context"), ([a, b], "two contexts")):
    run(ctxs, 4)
    t = time.perf_counter()
    run(ctxs, 12)
    dt = (time.perf_counter() - t) / 12 * 1e3
    print("%s: %.2f ms per MSM (%.1f Mpoints/s)" % (name, dt, n / dt / 1e3))
ref = [o.cpu().numpy().tobytes() for o in outs]
run([a], 2)
assert ref == [o.cpu().numpy().tobytes() for o in outs]
