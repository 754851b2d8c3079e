"""python scripts/msm_phases.py LOG2N [c]: phase breakdown of one MSM size in classic and window-table mode."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images
lg = int(sys.argv[1]); n = 1 << lg
ctx = zk.Context(0)
srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([12345678901234567890]), ctx)
sc = torch.from_numpy(images(n, 7)).cuda(); torch.cuda.synchronize()
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
for mode in ("classic", "table"):
    if mode == "table":
        srs.precompute(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for _ in range(2): zk.MultiExp(srs, sc, n=n, out=out)
    ctx.profile(True); ctx.profile_read()
    reps = 5
    for _ in range(reps): zk.MultiExp(srs, sc, n=n, out=out)
    ph = ctx.profile_read(); ctx.profile(False)
    print(mode, lg, {k: round(v[0] / reps, 3) for k, v in ph.items() if v[1]}, "total", round(sum(v[0] for v in ph.values()) / reps, 3))
