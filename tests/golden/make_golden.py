"""Generates tests/golden/*.json from the Python big-int oracle (oracle/bn254.py).

The reference holds no golden vectors for this path and cannot be run here (Go), so these
fixtures pin OUR restatement: they make the C oracle, the Python oracle and the CUDA path agree on fixed bytes
across rounds.  Regenerate with:  python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bn254 as o  # noqa: E402


def main():
    ntt = []
    for log2n in (0, 1, 2, 3, 5, 8):
        n = 1 << log2n
        a = o.random_fr(n, 0xB2000003 + log2n)
        d = o.Domain(n)
        for inverse in (0, 1):
            for decimation in (o.DIF, o.DIT):
                for coset in (0, 1):
                    out = (d.fft_inverse if inverse else d.fft)(a, decimation, bool(coset))
                    ntt.append({
                        "log2n": log2n, "inverse": inverse, "decimation": decimation, "coset": coset,
                        "in": o.fr_to_mont_bytes(a).hex(), "out": o.fr_to_mont_bytes(out).hex(),
                    })
    with open(os.path.join(HERE, "ntt_small.json"), "w") as f:
        json.dump(ntt, f)

    msm = []
    for n, seed in ((1, 1), (2, 2), (7, 3), (64, 4), (300, 5)):
        pts = o.g1_structured_bases(n, 0xB2000002 % 1000 + seed, 17 + seed)
        sc = o.random_fr(n, 0xB2000001 + seed)
        if n >= 7:
            pts[3] = None          # point at infinity among the bases
            sc[1] = 0              # zero scalar
            sc[2] = o.R_MOD - 1    # -1
            sc[5] = 1
        res = o.g1_msm(pts, sc, 6) if n > 64 else o.g1_msm_naive(pts, sc)
        msm.append({"n": n, "points": o.g1_to_bytes(pts).hex(), "scalars": o.fr_to_mont_bytes(sc).hex(),
                    "out": o.g1_to_bytes([res]).hex()})
    with open(os.path.join(HERE, "msm_small.json"), "w") as f:
        json.dump(msm, f)


if __name__ == "__main__":
    main()
