"""python scripts/fieldmul_probe.py: field-multiplier microbenchmarks (b200zk_microbench 0..7), the 2^24 MSM phases and
the 2^24 NTT with the library as built — run once per build variant to fill profiles/r02_fieldmul.md."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images
ctx = zk.Context(0)
names = ["imad_wide", "fp_mul(lib)", "fr_mul(lib)", "dfma", "imad_wide_beside_dfma", "fp_mul_schoolbook", "fp_mul_karatsuba",
         "fp_products_in_mul2add"]
print({names[k]: round(ctx.microbench(k) / 1e9, 2) for k in range(8)}, "G/s")
ext = ctx.torch_stream()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([12345678901234567890]), ctx).precompute()
sc = torch.from_numpy(images(n, 7)).cuda(); torch.cuda.synchronize()
out = torch.zeros(64, dtype=torch.uint8, device="cuda")
for mode in (0,):
    for _ in range(2): zk.MultiExp(srs, sc, n=n, out=out)
    ctx.profile(True); ctx.profile_read()
    for _ in range(5): zk.MultiExp(srs, sc, n=n, out=out)
    ph = ctx.profile_read(); ctx.profile(False)
    print("msm 2^%d table, front end %d:" % (lg, mode), {k: round(v[0] / 5, 3) for k, v in ph.items() if v[1]}, "sum", round(sum(v[0] for v in ph.values()) / 5, 3))
srs.close(); del sc
a = torch.from_numpy(images(n, 3)).cuda(); torch.cuda.synchronize()
d = zk.Domain(n, ctx)
for _ in range(3): d.FFT(a, zk.DIF, False)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ext)
for _ in range(10): d.FFT(a, zk.DIF, False)
e1.record(ext); ctx.sync()
print("ntt 2^%d DIF: %.3f ms" % (lg, e0.elapsed_time(e1) / 10))
