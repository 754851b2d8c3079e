"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the BN254 PLONK hot path: fr / fp arithmetic, G1, MSM, NTT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It is never on the product path: noir_backend_using_gnark_b200/ does not import it.

PARITY STATUS: **parity unpinned** against the real reference.  The arithmetic on this path lives in the
un-vendored Go modules github.com/consensys/gnark-crypto v0.9.1 and github.com/consensys/gnark v0.8.0
(/root/reference/gnark_backend_ffi/go.mod:5,23); neither a Go toolchain nor the module sources exist in this
environment and the reference's own tests hold no golden MSM / NTT / proof vectors.
This file therefore restates the *published* algorithms and conventions of those modules with Python big
integers and is pinned by independently computed known-answer vectors (tests/test_oracle_kat.py):
2G, 3G, r*G = inf, omega_4, the size-4 NTT of [1,2,3,4], BN parameter identities, and the one field constant
the reference itself pins (r-1 = 0x3064...0000 used as the coefficient "-1" at
/root/reference/gnark_backend_ffi/main.go:233).

Conventions restated (gnark-crypto v0.9.1):
  * fr.Element / fp.Element: 4 x u64 little-endian limbs, Montgomery form, R = 2^256   (ecc/bn254/fr/element.go)
  * G1Affine: X || Y (two fp.Element), point at infinity = (0, 0)                        (ecc/bn254/g1.go)
  * fft.Domain: Generator = g^(2^(28-log2 n)), g of order 2^28; FrMultiplicativeGen = 5;
    DIF: natural in -> bit-reversed out; DIT: bit-reversed in -> natural out;
    FFT(coset): pre-scale by 5^i;  FFTInverse: post-scale by 1/n (and 5^-i if coset)     (ecc/bn254/fr/fft/fft.go)
  * MultiExp: sum_i s_i * P_i with s_i given in Montgomery form; result canonical affine (ecc/bn254/multiexp.go)
"""
from __future__ import annotations

import struct
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------------
# constants  ([COMPUTED]; checked again in tests/test_oracle_kat.py)
# --------------------------------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # fr modulus r
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # fp modulus p
BN_X = 4965661367192848881
MONT_R = 1 << 256
FR_ROOT_2_28 = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # order 2^28
FR_MAX_LOG2 = 28
FR_COSET_GEN = 5  # fft.Domain.FrMultiplicativeGen
G1_GEN = (1, 2)
CURVE_B = 3

FR_R = MONT_R % R_MOD
FR_R2 = (MONT_R * MONT_R) % R_MOD
FR_RINV = pow(MONT_R, -1, R_MOD)
FP_R = MONT_R % P_MOD
FP_R2 = (MONT_R * MONT_R) % P_MOD
FP_RINV = pow(MONT_R, -1, P_MOD)
FR_NINV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
FP_NINV64 = (-pow(P_MOD, -1, 1 << 64)) % (1 << 64)
FR_NINV32 = FR_NINV64 & 0xFFFFFFFF
FP_NINV32 = FP_NINV64 & 0xFFFFFFFF


# --------------------------------------------------------------------------------------------------
# byte layouts
# --------------------------------------------------------------------------------------------------
def limbs_le(x: int) -> bytes:
    """4 x u64 little-endian limbs (the in-memory layout of fr.Element / fp.Element)."""
    return x.to_bytes(32, "little")


def from_limbs_le(b: bytes) -> int:
    return int.from_bytes(b, "little")


def fr_to_mont_bytes(xs: Iterable[int]) -> bytes:
    return b"".join(limbs_le((x % R_MOD) * MONT_R % R_MOD) for x in xs)


def fr_from_mont_bytes(b: bytes) -> List[int]:
    assert len(b) % 32 == 0
    return [from_limbs_le(b[i : i + 32]) * FR_RINV % R_MOD for i in range(0, len(b), 32)]


def fp_to_mont_bytes(xs: Iterable[int]) -> bytes:
    return b"".join(limbs_le((x % P_MOD) * MONT_R % P_MOD) for x in xs)


def fp_from_mont_bytes(b: bytes) -> List[int]:
    return [from_limbs_le(b[i : i + 32]) * FP_RINV % P_MOD for i in range(0, len(b), 32)]


Affine = Optional[Tuple[int, int]]  # None = point at infinity


def g1_to_bytes(pts: Iterable[Affine]) -> bytes:
    """G1Affine in-memory layout: X||Y Montgomery limbs; infinity = 64 zero bytes."""
    out = []
    for pt in pts:
        if pt is None:
            out.append(b"\0" * 64)
        else:
            out.append(fp_to_mont_bytes(pt))
    return b"".join(out)


def g1_from_bytes(b: bytes) -> List[Affine]:
    out: List[Affine] = []
    for i in range(0, len(b), 64):
        x, y = fp_from_mont_bytes(b[i : i + 64])
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def fr_be_bytes(x: int) -> bytes:
    """fr.Element.Bytes(): 32-byte big-endian regular form (the FFI wire format,
    /root/reference/src/gnark_backend_wrapper/serialize.rs:33-47)."""
    return (x % R_MOD).to_bytes(32, "big")


# --------------------------------------------------------------------------------------------------
# G1 (y^2 = x^3 + 3 over fp), affine big-int arithmetic
# --------------------------------------------------------------------------------------------------
def g1_is_on_curve(pt: Affine) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - CURVE_B) % P_MOD == 0


def g1_neg(pt: Affine) -> Affine:
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P_MOD)


def g1_add(a: Affine, b: Affine) -> Affine:
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    y3 = (lam * (x1 - x3) - y1) % P_MOD
    return (x3, y3)


# Jacobian for speed in scalar-mul / MSM (python ints)
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % P_MOD
    B = Y * Y % P_MOD
    C = B * B % P_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % P_MOD
    E = 3 * A % P_MOD
    F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % P_MOD
    Z2Z2 = Z2 * Z2 % P_MOD
    U1 = X1 * Z2Z2 % P_MOD
    U2 = X2 * Z1Z1 % P_MOD
    S1 = Y1 * Z2 * Z2Z2 % P_MOD
    S2 = Y2 * Z1 * Z1Z1 % P_MOD
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % P_MOD
    Rr = (S2 - S1) % P_MOD
    HH = H * H % P_MOD
    HHH = H * HH % P_MOD
    V = U1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - S1 * HHH) % P_MOD
    Z3 = Z1 * Z2 * H % P_MOD
    return (X3, Y3, Z3)


def _to_jac(pt: Affine):
    return (1, 1, 0) if pt is None else (pt[0], pt[1], 1)


def _from_jac(P) -> Affine:
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def g1_mul(pt: Affine, k: int) -> Affine:
    k %= R_MOD
    acc = (1, 1, 0)
    base = _to_jac(pt)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(base)
        k >>= 1
    return _from_jac(acc)


def g1_msm_naive(points: Sequence[Affine], scalars: Sequence[int]) -> Affine:
    """sum_i scalars[i]*points[i], straight definition of (*G1Affine).MultiExp (regular-form scalars)."""
    acc = (1, 1, 0)
    for pt, s in zip(points, scalars):
        acc = _jac_add(acc, _to_jac(g1_mul(pt, s)))
    return _from_jac(acc)


def g1_msm(points: Sequence[Affine], scalars: Sequence[int], c: int = 8) -> Affine:
    """Bucket-method MSM (unsigned c-bit windows) — same result as g1_msm_naive, usable to ~2^14 points."""
    n = len(points)
    assert len(scalars) == n
    nwin = (254 + c - 1) // c
    jp = [_to_jac(p) for p in points]
    total = (1, 1, 0)
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = _jac_double(total)
        buckets = [(1, 1, 0)] * (1 << c)
        for i in range(n):
            d = ((scalars[i] % R_MOD) >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d] = _jac_add(buckets[d], jp[i])
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            run = _jac_add(run, buckets[d])
            acc = _jac_add(acc, run)
        total = _jac_add(total, acc)
    return _from_jac(total)


def g1_structured_bases(n: int, a: int, b: int) -> List[Affine]:
    """P_i = (a + i*b)*G by repeated affine addition (SURVEY.md §8c self-check construction)."""
    step = g1_mul(G1_GEN, b)
    cur = g1_mul(G1_GEN, a)
    out = []
    for _ in range(n):
        out.append(cur)
        cur = g1_add(cur, step)
    return out


def g1_pseudo_random_points(n: int, seed: int) -> List[Affine]:
    """Bases with no known discrete-log relation (bases (ii)): x from a SplitMix64 stream, incremented until
    x^3 + 3 is a square; y = (x^3 + 3)^((p+1)/4) (p = 3 mod 4), the smaller root when the stream says so."""
    out = []
    state = seed & MASK64
    for _ in range(n):
        x = 0
        for _k in range(4):
            state, z = splitmix64(state)
            x = (x << 64) | z
        x %= P_MOD
        while True:
            rhs = (x * x * x + 3) % P_MOD
            y = pow(rhs, (P_MOD + 1) // 4, P_MOD)
            if y * y % P_MOD == rhs:
                break
            x = (x + 1) % P_MOD
        state, z = splitmix64(state)
        if z & 1:
            y = P_MOD - y
        out.append((x, y))
    return out


# --------------------------------------------------------------------------------------------------
# NTT with gnark fft.Domain semantics
# --------------------------------------------------------------------------------------------------
DIF = 0
DIT = 1


def bit_reverse_index(i: int, log2n: int) -> int:
    return int(format(i, "0%db" % log2n)[::-1], 2) if log2n else 0


def bit_reverse(a: List[int]) -> List[int]:
    """fft.BitReverse: in-place permutation a[i] <-> a[brev(i)]."""
    n = len(a)
    log2n = n.bit_length() - 1
    out = list(a)
    for i in range(n):
        out[bit_reverse_index(i, log2n)] = a[i]
    return out


class Domain:
    """Restatement of gnark-crypto v0.9.1 ecc/bn254/fr/fft.Domain (NewDomain + FFT / FFTInverse)."""

    def __init__(self, m: int):
        n = 1
        while n < m:
            n <<= 1
        self.cardinality = n
        self.log2n = n.bit_length() - 1
        assert self.log2n <= FR_MAX_LOG2
        self.generator = pow(FR_ROOT_2_28, 1 << (FR_MAX_LOG2 - self.log2n), R_MOD)
        self.generator_inv = pow(self.generator, -1, R_MOD)
        self.cardinality_inv = pow(n, -1, R_MOD)
        self.fr_multiplicative_gen = FR_COSET_GEN
        self.fr_multiplicative_gen_inv = pow(FR_COSET_GEN, -1, R_MOD)

    # -- kernels (radix-2, in the same loop nest order as gnark's difFFT / ditFFT) -----------------
    @staticmethod
    def _dif(a: List[int], w: int) -> None:
        n = len(a)
        m = n
        wm = w
        while m > 1:
            half = m >> 1
            tw = [1] * half
            for j in range(1, half):
                tw[j] = tw[j - 1] * wm % R_MOD
            for start in range(0, n, m):
                for j in range(half):
                    u = a[start + j]
                    v = a[start + j + half]
                    a[start + j] = (u + v) % R_MOD
                    a[start + j + half] = (u - v) * tw[j] % R_MOD
            wm = wm * wm % R_MOD
            m = half

    @staticmethod
    def _dit(a: List[int], w: int) -> None:
        n = len(a)
        log2n = n.bit_length() - 1
        m = 2
        while m <= n:
            half = m >> 1
            wm = pow(w, n // m, R_MOD)
            tw = [1] * half
            for j in range(1, half):
                tw[j] = tw[j - 1] * wm % R_MOD
            for start in range(0, n, m):
                for j in range(half):
                    u = a[start + j]
                    v = a[start + j + half] * tw[j] % R_MOD
                    a[start + j] = (u + v) % R_MOD
                    a[start + j + half] = (u - v) % R_MOD
            m <<= 1
        del log2n

    def fft(self, a: List[int], decimation: int, coset: bool = False) -> List[int]:
        """domain.FFT(a, decimation, coset): returns the transformed copy (gnark works in place)."""
        n = self.cardinality
        assert len(a) == n
        a = [x % R_MOD for x in a]
        if coset:
            # CosetTable[i] = 5^i; DIT input is bit-reversed so the table is used bit-reversed
            g = self.fr_multiplicative_gen
            tab = [1] * n
            for i in range(1, n):
                tab[i] = tab[i - 1] * g % R_MOD
            if decimation == DIT:
                tab = bit_reverse(tab)
            a = [x * t % R_MOD for x, t in zip(a, tab)]
        if decimation == DIF:
            self._dif(a, self.generator)
        else:
            self._dit(a, self.generator)
        return a

    def fft_inverse(self, a: List[int], decimation: int, coset: bool = False) -> List[int]:
        """domain.FFTInverse(a, decimation, coset)."""
        n = self.cardinality
        assert len(a) == n
        a = [x % R_MOD for x in a]
        if decimation == DIF:
            self._dif(a, self.generator_inv)
        else:
            self._dit(a, self.generator_inv)
        ninv = self.cardinality_inv
        if not coset:
            return [x * ninv % R_MOD for x in a]
        gi = self.fr_multiplicative_gen_inv
        tab = [1] * n
        for i in range(1, n):
            tab[i] = tab[i - 1] * gi % R_MOD
        if decimation == DIF:  # output is bit-reversed -> CosetTableInvReversed
            tab = bit_reverse(tab)
        return [x * t % R_MOD * ninv % R_MOD for x, t in zip(a, tab)]


def ntt_naive(a: Sequence[int], w: int) -> List[int]:
    """O(n^2) definition: X[k] = sum_j a[j] w^(jk) (natural order in and out)."""
    n = len(a)
    return [sum(a[j] * pow(w, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]


# --------------------------------------------------------------------------------------------------
# seeded synthetic inputs shared by oracle, tests and bench (SplitMix64; BASELINE.md seeds)
# --------------------------------------------------------------------------------------------------
MASK64 = (1 << 64) - 1


def splitmix64(state: int) -> Tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & MASK64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return state, z ^ (z >> 31)


def random_fr(n: int, seed: int) -> List[int]:
    """Uniform elements of [0, r): 4 SplitMix64 words, top 2 bits cleared, rejection-sampled."""
    out = []
    st = seed & MASK64
    while len(out) < n:
        v = 0
        for k in range(4):
            st, z = splitmix64(st)
            v |= z << (64 * k)
        v &= (1 << 254) - 1
        if v < R_MOD:
            out.append(v)
    return out


def pack_u64(words: Sequence[int]) -> bytes:
    return struct.pack("<%dQ" % len(words), *words)
