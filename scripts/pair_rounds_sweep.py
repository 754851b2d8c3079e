"""python scripts/pair_rounds_sweep.py LOG2N [LOG2N ...]: MSM time (window table, then classic windows) per number of
batched-affine pair rounds (b200zk_msm_set_pair_rounds) and per batch length (B200ZK_MSM_PAIR_KMAX is read at context
creation, so the batch length is swept through separate contexts).  Every setting must return the same bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from sweep import images

sizes = [int(a) for a in sys.argv[1:]] or [24]
kmaxes = [int(k) for k in os.environ.get("KMAXES", "512").split(",")]
modes = os.environ.get("MODES", "table,classic").split(",")
for lg in sizes:
    n = 1 << lg
    for kmax in kmaxes:
        os.environ["B200ZK_MSM_PAIR_KMAX"] = str(kmax)
        ctx = zk.Context(0)
        lib = zk.load()
        srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([12345678901234567890]), ctx)
        sc = torch.from_numpy(images(n, 7)).cuda(); torch.cuda.synchronize()
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        for mode in ("classic", "table"):
            if mode == "table":
                srs.precompute()
            if mode not in modes:
                continue
            ref = None
            for rounds in (0, 1, 2, 3, 4, 5, -1):
                lib.b200zk_msm_set_pair_rounds(ctx.handle, rounds)
                for _ in range(2): zk.MultiExp(srs, sc, n=n, out=out)
                ctx.sync()
                got = bytes(out.cpu().numpy())
                ref = ref or got
                ctx.profile(True); ctx.profile_read()
                reps = 5
                for _ in range(reps): zk.MultiExp(srs, sc, n=n, out=out)
                ph = ctx.profile_read(); ctx.profile(False)
                print(json.dumps({"log2n": lg, "mode": mode, "kmax": kmax, "rounds": rounds, "same": got == ref,
                                  "ms": {k: round(v[0] / reps, 3) for k, v in ph.items() if v[1]},
                                  "total": round(sum(v[0] for v in ph.values()) / reps, 3)}), flush=True)
        srs.close(); del sc
        ctx.close() if hasattr(ctx, "close") else None
