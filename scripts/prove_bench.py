"""PLONK prove latency on one GPU: python scripts/prove_bench.py LOG2N [REPS] [--verify]
Synthetic chain circuit x_{i+1} = x_i^2 + x_i + c_i with 2^LOG2N - 1 gates + 1 public input."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp

R = zkp.R_MOD


def synthetic(log2n: int, seed: int = 0xB2000004):
    n = 1 << log2n
    gates = n - 1
    rng = np.random.default_rng(seed)
    limbs = rng.integers(0, 1 << 63, size=(gates, 4), dtype=np.uint64)
    limbs[:, 3] &= (1 << 60) - 1                    # < r: any residue is a valid Montgomery image
    qk_mont = limbs.view(np.uint8).reshape(-1)       # c_i in Montgomery form
    rinv = pow(1 << 256, -1, R)
    cb = qk_mont.tobytes()
    x = [int(rng.integers(1, 1 << 62))]
    for i in range(gates):
        c = int.from_bytes(cb[32 * i:32 * i + 32], "little") * rinv % R
        x.append((x[i] * x[i] + x[i] + c) % R)
    one = zkp.fr_to_mont([1])
    mone = zkp.fr_to_mont([R - 1])
    zero = np.zeros(32, dtype=np.uint8)
    ql = np.concatenate([mone, np.tile(one, gates)])            # placeholder row: -1
    qm = np.concatenate([zero, np.tile(one, gates)])
    qo = np.concatenate([zero, np.tile(mone, gates)])
    qr = np.zeros(n * 32, dtype=np.uint8)
    qk = np.concatenate([zero, qk_mont])
    lro = np.zeros(3 * n, dtype=np.uint32)
    idx = np.arange(gates, dtype=np.uint32)
    lro[0] = 0
    lro[1:n] = idx
    lro[n + 1:2 * n] = idx
    lro[2 * n + 1:3 * n] = idx + 1
    sol = zkp.fr_to_mont(x)
    return dict(ql=ql, qr=qr, qm=qm, qo=qo, qk=qk, lro=lro, sol=sol, x0=x[0], nb_wires=n)


def main():
    log2n = int(sys.argv[1])
    reps = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 3
    ctx = zk.Context(0)
    n = 1 << log2n
    t0 = time.time()
    c = synthetic(log2n)
    t_syn = time.time() - t0
    alpha = zkp.fr_to_mont([0x1234567890ABCDEF1234567])
    t0 = time.time()
    srs = zk.SRS.NewSRS(n + 3, alpha, ctx)
    if "--no-precompute" not in sys.argv:
        srs.precompute()
    t_srs = time.time() - t0
    t0 = time.time()
    pk = zkp.ProvingKey.SetupRaw(srs, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"], c["lro"], ctx)
    t_setup = time.time() - t0
    blind = np.frombuffer(os.urandom(9 * 32), dtype=np.uint8).copy()
    blind[31::32] &= 0x0F
    proof = pk.Prove(c["sol"], blind)       # warm-up (builds domains, workspaces)
    ctx.profile(True)
    ctx.profile_read()
    l0 = ctx.launch_count
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        proof = pk.Prove(c["sol"], blind)
        times.append((time.perf_counter() - t0) * 1e3)
    phases = ctx.profile_read()
    launches = (ctx.launch_count - l0) // reps
    ctx.profile(False)
    res = {"log2n": log2n, "prove_ms_best": min(times), "prove_ms_all": times, "setup_s": t_setup, "srs_s": t_srs,
           "synth_s": t_syn, "launches_per_prove": launches,
           "phase_ms_per_prove": {k: v[0] / reps for k, v in phases.items() if v[1]}}
    if "--verify" in sys.argv:
        from oracle import bn254 as o
        from oracle import plonk as pl
        a = 0x1234567890ABCDEF1234567
        S = [o.g1_from_bytes(b)[0] for b in pk.vk_points]
        vk = pl.VerifyingKey(n, pow(n, -1, R), o.Domain(n).generator, 1, 5, S[:3], S[3], S[4], S[5], S[6], S[7])
        g2 = (pl.G2_GEN, pl.g2_mul(pl.G2_GEN, a))
        res["verified"] = bool(pl.verify(pl.Proof.from_bytes(proof.to_gnark_bytes()), vk, [c["x0"]], g2))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
