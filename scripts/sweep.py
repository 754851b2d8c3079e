"""Size sweeps of BASELINE.json configs 4 and 5 on one GPU:
     python scripts/sweep.py msm [LO HI]     G1 MSM 2^LO..2^HI random scalars vs device SRS (classic + window table)
     python scripts/sweep.py ntt [LO HI]     fr NTT 2^LO..2^HI, {FFT, FFTInverse} x {DIF, DIT} x {plain, coset}
   Every size is self-checked without the oracle: MSM against the closed form Commit(p) = p(alpha)*G (one scalar
   multiplication through the same library at n = 1) and table mode against classic mode; NTT by round trips.
   Prints one JSON object per size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp

R = zkp.R_MOD


def images(n, seed):
    rng = np.random.default_rng(seed)
    limbs = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    limbs[:, 3] &= (1 << 60) - 1
    return limbs.view(np.uint8).reshape(-1)


def timeit(ctx, fn, reps):
    ext = ctx.torch_stream()
    for _ in range(2):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(reps):
        fn()
    e1.record(ext)
    ctx.sync()
    return e0.elapsed_time(e1) / reps


def msm_sweep(ctx, lo, hi):
    alpha = 0x1234567890ABCDEF1234567890ABCDEF % R
    for lg in range(lo, hi + 1):
        n = 1 << lg
        srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([alpha]), ctx)
        sc = torch.from_numpy(images(n, 0xB2000001 + lg)).cuda()
        torch.cuda.synchronize()
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        reps = 10 if lg <= 22 else 4
        t_classic = timeit(ctx, lambda: zk.MultiExp(srs, sc, n=n, out=out), reps)
        res_classic = out.cpu().numpy().tobytes()
        srs.precompute()
        t_table = timeit(ctx, lambda: zk.MultiExp(srs, sc, n=n, out=out), reps)
        res_table = out.cpu().numpy().tobytes()
        ok = res_classic == res_table
        if lg <= 20:  # closed form p(alpha)*G: Horner on the host in Python ints (slow beyond 2^20)
            vals = zkp.fr_from_mont(sc.cpu().numpy().tobytes())
            acc = 0
            for v in reversed(vals):
                acc = (acc * alpha + v) % R
            g = zk.SRS.NewSRS(1, zkp.fr_to_mont([1]), ctx)
            want = zk.MultiExp(g, zkp.fr_to_mont([acc]))
            ok = ok and want == res_table
            g.close()
        print(json.dumps({"op": "msm", "log2n": lg, "classic_ms": t_classic, "table_ms": t_table,
                          "classic_mpts": n / t_classic / 1e3, "table_mpts": n / t_table / 1e3, "self_check": ok}), flush=True)
        srs.close()
        del sc


def ntt_sweep(ctx, lo, hi):
    for lg in range(lo, hi + 1):
        n = 1 << lg
        x = torch.from_numpy(images(n, 0xB2000003 + lg)).cuda()
        torch.cuda.synchronize()
        d = zk.Domain(n, ctx)
        reps = 20 if lg <= 22 else 5
        row = {"op": "ntt", "log2n": lg}
        for name, inv, dec, cos in (("fft_dif", 0, zk.DIF, False), ("fft_dif_coset", 0, zk.DIF, True),
                                    ("ifft_dit", 1, zk.DIT, False), ("ifft_dit_coset", 1, zk.DIT, True),
                                    ("fft_dit", 0, zk.DIT, False), ("ifft_dif_coset", 1, zk.DIF, True)):
            y = x.clone()
            torch.cuda.synchronize()
            f = (lambda: d.FFTInverse(y, dec, cos)) if inv else (lambda: d.FFT(y, dec, cos))
            ms = timeit(ctx, f, reps)
            row[name + "_ms"] = ms
            row[name + "_gbs"] = 64.0 * n / ms / 1e6
            del y
        # round trips: FFT(DIF) o FFTInverse(DIT) = id, plain and coset
        ok = True
        for cos in (False, True):
            y = x.clone()
            torch.cuda.synchronize()
            d.FFT(y, zk.DIF, cos)
            d.FFTInverse(y, zk.DIT, cos)
            ctx.sync()
            ok = ok and bool(torch.equal(y, x))
            del y
        row["roundtrip_ok"] = ok
        print(json.dumps(row), flush=True)
        del x


def skew_sweep(ctx, lg):
    """adversarial scalar sets of SURVEY.md 8d at one size: all-equal, witness-like (50 % zero, 25 % < 2^16, 25 % uniform)"""
    n = 1 << lg
    alpha = 0x1234567890ABCDEF1234567890ABCDEF % R
    srs = zk.SRS.NewSRS(n, zkp.fr_to_mont([alpha]), ctx).precompute()
    uni = images(n, 0xB2000001)
    one = uni[:32]
    sets = {"uniform": uni, "all_equal": np.tile(one, n)}
    w = uni.copy().reshape(n, 32)
    w[: n // 2] = 0
    small = np.zeros((n // 4, 32), dtype=np.uint8)
    small[:, :2] = uni.reshape(n, 32)[: n // 4, :2]
    # small values must be stored in Montgomery form to be small scalars: convert through the library-free path
    vals = [int.from_bytes(bytes(r[:2]), "little") for r in small]
    w[n // 2: n // 2 + n // 4] = zkp.fr_to_mont(vals).reshape(-1, 32)
    sets["witness_like"] = w.reshape(-1)
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    for name, arr in sets.items():
        sc = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        torch.cuda.synchronize()
        t = timeit(ctx, lambda: zk.MultiExp(srs, sc, n=n, out=out), 3)
        print(json.dumps({"op": "msm_skew", "log2n": lg, "set": name, "table_ms": t, "mpts": n / t / 1e3}), flush=True)
    srs.close()


if __name__ == "__main__":
    what = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else (26 if what == "msm" else 28)
    ctx = zk.Context(0)
    if what == "skew":
        skew_sweep(ctx, lo)
    else:
        (msm_sweep if what == "msm" else ntt_sweep)(ctx, lo, hi)
