"""Tiny driver for ncu captures: python scripts/prof.py {msm|ntt} LOG2N REPS [c]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk

what, log2n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = zk.Context(0)
n = 1 << log2n
rng = np.random.default_rng(1)
# uniform 254-bit values below r are fine as Montgomery images (any residue is a valid element)
raw = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
raw[:, 3] &= (1 << 60) - 1
x = torch.from_numpy(raw.view(np.uint8).reshape(-1)).cuda()
torch.cuda.synchronize()
if what == "msm":
    if len(sys.argv) > 4:
        zk.load().b200zk_msm_set_window(ctx.handle, int(sys.argv[4]))
    one = (1).to_bytes(32, "little")
    srs = zk.SRS.NewSRS(n, bytes(raw[0].tobytes()), ctx)
    if os.environ.get("B200ZK_PRECOMPUTE", "1") == "1" and len(sys.argv) <= 4:
        srs.precompute()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        zk.MultiExp(srs, x, n=n, out=out)
    ctx.sync()
else:
    d = zk.Domain(n, ctx)
    for _ in range(reps):
        d.FFT(x, zk.DIF, False)
    ctx.sync()
print("done")
