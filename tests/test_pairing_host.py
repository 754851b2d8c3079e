"""CPU test of the product's host pairing (csrc/ffi/bn254_host.h: Miller loop on the twist, Frobenius by constants,
final exponentiation split into easy part and the p-adic expansion of the hard part) against the oracle's big-int
pairing, which works generically in Fp12 and raises to (p^12 - 1)/r bit by bit.  Both use Fp[w]/(w^12 - 18 w^6 + 82),
so the values must agree coefficient by coefficient."""
import os
import subprocess

import pytest

from oracle import bn254 as o
from oracle import plonk as pl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = tmp_path_factory.mktemp("bin") / "pairing_dump"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "pairing_dump.cpp")], check=True)
    return str(exe)


@pytest.mark.parametrize("a,b", [(1, 1), (0x1234567, 0x89ABCDEF01), (0xFFFFFFFFFFFFFFFF, 3)])
def test_pairing_value_matches_the_oracle(dumper, a, b):
    out = subprocess.run([dumper, hex(a), hex(b)], capture_output=True, text=True, check=True).stdout.splitlines()
    P, Q = o.g1_mul(o.G1_GEN, a), pl.g2_mul(pl.G2_GEN, b)
    ml = pl.miller_loop(Q, P)
    assert out[0].split() == ["%064x" % c for c in ml]
    assert out[1].split() == ["%064x" % c for c in pl.final_exponentiation(ml)]
    ok, bad, ms = out[2].split()
    assert (ok, bad) == ("1", "0")                      # e(aG, bH) e(-abG, H) == 1, and != 1 for ab + 1
    assert float(ms) < 500
    in_g2, off_g2, _k = out[3].split()
    assert (in_g2, off_g2) == ("1", "0")                # [r]Q = O for Q in G2, not for a point of the twist outside it


def test_hard_part_expansion_is_exact():
    # the identity final_exponentiation() relies on (bn254_host.h), over the integers
    p, r, x = o.P_MOD, o.R_MOD, 4965661367192848881
    lam = [-(36 * x**3 + 30 * x**2 + 18 * x + 2), -(36 * x**3 + 18 * x**2 + 12 * x - 1), 6 * x * x + 1, 1]
    assert sum(l * p**i for i, l in enumerate(lam)) * r == p**4 - p**2 + 1
    assert (p**12 - 1) // r == (p**6 - 1) * (p**2 + 1) * ((p**4 - p**2 + 1) // r)
    assert max(abs(l) for l in lam).bit_length() <= 192
