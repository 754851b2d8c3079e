"""python scripts/msm_shard_phases.py LOG2_BASES: phase breakdown of the MSMs a rank of the sharded prover runs — a
2^LOG2_BASES-point SRS with its window table, MSMs over 1/1, 1/2, 1/4, 1/8 of the points (CUDA events per phase)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images
lg = int(sys.argv[1]); n = 1 << lg
ctx = zk.Context(0)
srs = zk.SRS.NewSRS(n + 3, zkp.fr_to_mont([12345678901234567890]), ctx).precompute()
sc = torch.from_numpy(images(n, 7)).cuda(); torch.cuda.synchronize()
out = torch.zeros(128, dtype=torch.uint8, device="cuda")
ext = ctx.torch_stream()
for g in (1, 2, 4, 8):
    m = n // g
    for _ in range(2): zk.MultiExp(srs, sc[: m * 32], n=m, first_base=m if g > 1 else 0, out=out, partial=True)
    ctx.profile(True); ctx.profile_read()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(reps): zk.MultiExp(srs, sc[: m * 32], n=m, first_base=m if g > 1 else 0, out=out, partial=True)
    e1.record(ext)
    ph = ctx.profile_read(); ctx.profile(False)
    print("bases 2^%d windows %d  msm over 2^%d points:" % (lg, srs.windows(m), lg - g.bit_length() + 1),
          {k: round(v[0] / reps, 3) for k, v in ph.items() if v[1]}, "sum", round(sum(v[0] for v in ph.values()) / reps, 3),
          "wall", round(e0.elapsed_time(e1) / reps, 3))
