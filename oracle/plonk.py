"""CPU oracle for the PLONK flow the reference drives (TEST INFRASTRUCTURE ONLY; parity unpinned — see below).

Restates, with Python big integers (NTT / MSM delegated to the C oracle for speed):
  * the reference's own glue [REF]:  ACIR JSON -> SparseR1CS
      /root/reference/gnark_backend_ffi/acir/**                       (JSON shapes)
      /root/reference/gnark_backend_ffi/backend/common.go:45-76       (HandleValues: public/secret partition)
      /root/reference/gnark_backend_ffi/backend/plonk/sparse_r1cs.go:44-107 (handleArithmeticOpcode, incl. its quirks)
  * gnark v0.8.0 backend/plonk/bn254 Setup / Prove / Verify and gnark-crypto v0.9.1 kzg / fiat-shamir, as called at
      /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:21 (Setup), :67 (Prove), :47 (Verify).
    Those sources are NOT under /root/reference (go.mod:5,23) and cannot be fetched or built here, so this part is
    written from the published algorithm.  **Parity unpinned**: byte identity with
    real gnark proofs cannot be established in this environment; what IS established is (a) internal consistency —
    every proof passes verify() below, which implements the verifier equations with an independent optimal-ate
    pairing, and (b) the CUDA prover reproduces this oracle's proof bytes exactly for the same blinding stream.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import hashlib
import json
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import bn254 as o
from . import cref

R = o.R_MOD
P = o.P_MOD
Affine = o.Affine


# ======================================================================================================
# 1. ACIR -> SparseR1CS   (reference glue, [REF])
# ======================================================================================================
@dataclass
class Gate:
    """constraint.SparseR1C: qL*xa + qR*xb + qO*xc + qM*(xa*xb) + qC == 0 (sparse_r1cs.go:17)."""
    ql: int
    qr: int
    qm: int
    qo: int
    qk: int
    a: int  # L.WireID
    b: int  # R.WireID
    c: int  # O.WireID


@dataclass
class SparseR1CS:
    nb_public: int = 0
    nb_secret: int = 0
    gates: List[Gate] = field(default_factory=list)


def decode_acir(acir_json: str) -> dict:
    """acir.ACIR.UnmarshalJSON (acir/acir.go:17-75): keys opcodes, public_inputs, current_witness_index; an opcode is
    {"Arithmetic": {"mul_terms": [[hex,w,w]], "linear_combinations": [[hex,w]], "q_c": hex}} | BlackBoxFuncCall |
    Directive (opcode/opcode.go:13-36)."""
    d = json.loads(acir_json)
    ops = []
    for op in d["opcodes"]:
        if "Arithmetic" in op:
            a = op["Arithmetic"]
            ops.append(("arith",
                        [(int(t[0], 16) % R, int(t[1]), int(t[2])) for t in a["mul_terms"]],
                        [(int(t[0], 16) % R, int(t[1])) for t in a["linear_combinations"]],
                        int(a["q_c"], 16) % R))
        elif "BlackBoxFuncCall" in op:
            ops.append(("blackbox",))      # components.go:3-40: no constraints
        elif "Directive" in op:
            ops.append(("directive",))     # sparse_r1cs.go:36: skipped
        else:
            raise ValueError("unknown opcode type")
    return {"current_witness": int(d["current_witness_index"]), "opcodes": ops,
            "public_inputs": [int(x) for x in d["public_inputs"]]}


def handle_values(acir: dict, values: Sequence[int]):
    """backend.HandleValues (common.go:45-76), bug-compatible: with P public inputs every non-public value is
    registered P times as a secret (and public values P-1 extra times); index_map keeps the last registration."""
    public_vals, secret_vals, index_map = [], [], {}
    nb_public = nb_secret = 0
    pubs = acir["public_inputs"]
    for i, v in enumerate(values, start=1):
        for p in pubs:
            if i == p:
                index_map[i] = nb_public           # AddPublicVariable -> index among publics
                nb_public += 1
                public_vals.append(v)
    for i, v in enumerate(values, start=1):
        if pubs:
            for p in pubs:
                if i != p:
                    index_map[i] = nb_public + nb_secret   # AddSecretVariable -> len(Public)+len(Secret)
                    nb_secret += 1
                    secret_vals.append(v)
        else:
            index_map[i] = nb_public + nb_secret
            nb_secret += 1
            secret_vals.append(v)
    return public_vals, secret_vals, index_map, nb_public, nb_secret


def build_sparse_r1cs(acir: dict, values: Sequence[int]):
    """plonk_backend.BuildSparseR1CS (sparse_r1cs.go:18-25, 44-107)."""
    pub, sec, imap, nb_public, nb_secret = handle_values(acir, values)
    cs = SparseR1CS(nb_public, nb_secret)
    for op in acir["opcodes"]:
        if op[0] != "arith":
            continue
        _, mul_terms, lin, qc = op
        xa = xb = xc = 0
        ql = qr = qo = qm1 = 0
        qm2 = 0
        if mul_terms:                       # only MulTerms[0] is read (:50)
            qm1, wa, wb = mul_terms[0]
            qm2 = 1
            xa, xb = imap.get(wa, 0), imap.get(wb, 0)
        if len(lin) == 1:
            qo, w = lin[0]
            xc = imap.get(w, 0)
        if len(lin) == 2:
            (ql, w0), (qr, w1) = lin
            xa, xb = imap.get(w0, 0), imap.get(w1, 0)   # overwrites the mul term's wires (:69,:73)
        if len(lin) == 3:
            (ql, w0), (qr, w1), (qo, w2) = lin
            xa, xb, xc = imap.get(w0, 0), imap.get(w1, 0), imap.get(w2, 0)
        cs.gates.append(Gate(ql, qr, qm1 * qm2 % R, qo, qc, xa, xb, xc))
    return cs, pub, sec


# ======================================================================================================
# 2. helpers: NTT / MSM through the C oracle, encodings, transcript
# ======================================================================================================
def _fft(vals: List[int], inverse: bool, decimation: int, coset: bool) -> List[int]:
    n = len(vals)
    log2n = n.bit_length() - 1
    out = cref.ntt(o.fr_to_mont_bytes(vals), log2n, inverse, decimation, coset, nthreads=cref.ncores())
    return o.fr_from_mont_bytes(out)


def to_canonical(lagrange: List[int]) -> List[int]:
    """FFTInverse(DIF) + BitReverse on Domain[0] (Lagrange regular -> canonical regular)."""
    return o.bit_reverse(_fft(lagrange, True, o.DIF, False))


def to_lagrange_coset_bitrev(canonical: List[int], size: int) -> List[int]:
    """iop ToLagrangeCoset on Domain[1]: zero-pad, FFT(DIF, coset) -> evaluations on 5*<w_4n>, bit-reversed layout."""
    return _fft(list(canonical) + [0] * (size - len(canonical)), False, o.DIF, True)


class SRS:
    """kzg.SRS: G1 = [alpha^i G1], G2 = [G2, alpha G2] (kzg.NewSRS, common.go:137 / main.go:176)."""

    def __init__(self, size: int, alpha: int):
        self.alpha = alpha % R
        pw = [1] * size
        for i in range(1, size):
            pw[i] = pw[i - 1] * self.alpha % R
        self.g1_bytes = cref.g1_mul_gen_batch(o.fr_to_mont_bytes(pw))
        self.size = size
        self.g2 = (G2_GEN, g2_mul(G2_GEN, self.alpha))


def commit(poly: Sequence[int], srs: SRS) -> Affine:
    """kzg.Commit: MultiExp(srs.G1[:len(p)], p)."""
    assert len(poly) <= srs.size, "SRS too small"
    out = cref.msm(srs.g1_bytes, o.fr_to_mont_bytes(poly), len(poly), nthreads=cref.ncores())
    return o.g1_from_bytes(out)[0]


def g1_marshal(pt: Affine) -> bytes:
    """G1Affine.Marshal() = RawBytes(): 64 B uncompressed, infinity = 0x40 || 0...."""
    if pt is None:
        return b"\x40" + b"\0" * 63
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g1_compress(pt: Affine) -> bytes:
    """G1Affine.Bytes(): 32 B big-endian X, top two bits: 10 = smallest y, 11 = largest y, 01 = infinity."""
    if pt is None:
        return b"\x40" + b"\0" * 31
    x, y = pt
    flag = 0x80 if y <= (P - 1) // 2 else 0xC0
    b = bytearray(x.to_bytes(32, "big"))
    b[0] |= flag
    return bytes(b)


def g1_decompress(b: bytes) -> Affine:
    flag = b[0] & 0xC0
    if flag == 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x3F]) + b[1:32], "big")
    y = pow((x * x * x + 3) % P, (P + 1) // 4, P)
    assert y * y % P == (x * x * x + 3) % P, "not on curve"
    if (y <= (P - 1) // 2) != (flag == 0x80):
        y = P - y
    return (x, y)


class Transcript:
    """fiatshamir.NewTranscript(sha256.New(), names...) (gnark-crypto v0.9.1 fiat-shamir/transcript.go)."""

    def __init__(self, *names: str):
        self.names = list(names)
        self.bindings: Dict[str, List[bytes]] = {n: [] for n in names}
        self.values: Dict[str, bytes] = {}

    def bind(self, name: str, data: bytes) -> None:
        assert name not in self.values
        self.bindings[name].append(bytes(data))

    def compute(self, name: str) -> bytes:
        if name in self.values:
            return self.values[name]
        h = hashlib.sha256()
        h.update(name.encode())
        pos = self.names.index(name)
        if pos != 0:
            h.update(self.values[self.names[pos - 1]])
        for b in self.bindings[name]:
            h.update(b)
        self.values[name] = h.digest()
        return self.values[name]


def challenge_to_fr(b: bytes) -> int:
    return int.from_bytes(b, "big") % R  # fr.Element.SetBytes


def derive_randomness(fs: Transcript, name: str, *points: Affine) -> int:
    for pt in points:
        fs.bind(name, g1_marshal(pt))
    return challenge_to_fr(fs.compute(name))


def eval_poly(p: Sequence[int], z: int) -> int:
    acc = 0
    for c in reversed(p):
        acc = (acc * z + c) % R
    return acc


def divide_by_x_minus_a(f: Sequence[int], fa: int, a: int) -> List[int]:
    """kzg.dividePolyByXminusA: (f - f(a)) / (X - a)."""
    f = list(f)
    f[0] = (f[0] - fa) % R
    for i in range(len(f) - 2, -1, -1):
        f[i] = (f[i] + f[i + 1] * a) % R
    return f[1:]


# ======================================================================================================
# 3. Setup (gnark v0.8.0 plonk.Setup)
# ======================================================================================================
@dataclass
class VerifyingKey:
    size: int
    size_inv: int
    generator: int
    nb_public: int
    coset_shift: int
    S: List[Affine]
    Ql: Affine
    Qr: Affine
    Qm: Affine
    Qo: Affine
    Qk: Affine


@dataclass
class ProvingKey:
    vk: VerifyingKey
    n: int
    n_big: int
    ql: List[int]
    qr: List[int]
    qm: List[int]
    qo: List[int]
    cqk: List[int]      # canonical, without public inputs
    lqk: List[int]      # Lagrange, to be completed by the prover
    s1: List[int]
    s2: List[int]
    s3: List[int]
    permutation: List[int]
    lro_wires: List[int]   # position -> wire id (3n), kept for the device prover
    # Lagrange-coset (bit-reversed) forms on Domain[1], recomputed at pk load in gnark
    l_ql: List[int] = field(default_factory=list)
    l_qr: List[int] = field(default_factory=list)
    l_qm: List[int] = field(default_factory=list)
    l_qo: List[int] = field(default_factory=list)
    l_s1: List[int] = field(default_factory=list)
    l_s2: List[int] = field(default_factory=list)
    l_s3: List[int] = field(default_factory=list)


def next_pow2(x: int) -> int:
    n = 1
    while n < x:
        n <<= 1
    return n


def setup(cs: SparseR1CS, srs: SRS) -> ProvingKey:
    nb_constraints = len(cs.gates)
    size_system = nb_constraints + cs.nb_public
    n = next_pow2(size_system)
    n_big = next_pow2(8 * size_system if size_system < 6 else 4 * size_system)
    dom = o.Domain(n)
    ql, qr, qm, qo, cqk, lqk = ([0] * n for _ in range(6))
    for i in range(cs.nb_public):
        ql[i] = R - 1                       # placeholder -PUB_i + qk_i = 0
    off = cs.nb_public
    for i, g in enumerate(cs.gates):
        ql[off + i], qr[off + i], qm[off + i], qo[off + i] = g.ql, g.qr, g.qm, g.qo
        cqk[off + i] = g.qk
        lqk[off + i] = g.qk
    ql_c, qr_c, qm_c, qo_c, cqk_c = (to_canonical(x) for x in (ql, qr, qm, qo, cqk))

    # buildPermutation
    lro = [0] * (3 * n)
    for i in range(cs.nb_public):
        lro[i] = i
    for i, g in enumerate(cs.gates):
        lro[off + i] = g.a
        lro[n + off + i] = g.b
        lro[2 * n + off + i] = g.c
    nb_vars = cs.nb_public + cs.nb_secret
    cycle = [-1] * max(nb_vars, 1)
    perm = [-1] * (3 * n)
    for i in range(3 * n):
        if cycle[lro[i]] != -1:
            perm[i] = cycle[lro[i]]
        cycle[lro[i]] = i
    for i in range(3 * n):
        if perm[i] == -1:
            perm[i] = cycle[lro[i]]

    # permutation polynomials
    ident = identity_support(dom)
    s1 = to_canonical([ident[perm[i]] for i in range(n)])
    s2 = to_canonical([ident[perm[n + i]] for i in range(n)])
    s3 = to_canonical([ident[perm[2 * n + i]] for i in range(n)])

    vk = VerifyingKey(n, pow(n, -1, R), dom.generator, cs.nb_public, o.FR_COSET_GEN,
                      [commit(s1, srs), commit(s2, srs), commit(s3, srs)],
                      commit(ql_c, srs), commit(qr_c, srs), commit(qm_c, srs), commit(qo_c, srs), commit(cqk_c, srs))
    pk = ProvingKey(vk, n, n_big, ql_c, qr_c, qm_c, qo_c, cqk_c, lqk, s1, s2, s3, perm, lro)
    # computeLagrangeCosetPolys (done in ReadFrom / Setup)
    pk.l_ql, pk.l_qr, pk.l_qm, pk.l_qo, pk.l_s1, pk.l_s2, pk.l_s3 = (
        to_lagrange_coset_bitrev(x, n_big) for x in (ql_c, qr_c, qm_c, qo_c, s1, s2, s3))
    return pk


def identity_support(dom: o.Domain) -> List[int]:
    """getIDSmallDomain: [w^i] || [u w^i] || [u^2 w^i]."""
    n = dom.cardinality
    u = o.FR_COSET_GEN
    res = [0] * (3 * n)
    res[0], res[n], res[2 * n] = 1, u, u * u % R
    for i in range(1, n):
        res[i] = res[i - 1] * dom.generator % R
        res[n + i] = res[n + i - 1] * dom.generator % R
        res[2 * n + i] = res[2 * n + i - 1] * dom.generator % R
    return res


# ======================================================================================================
# 4. Prove (gnark v0.8.0 plonk.Prove)
# ======================================================================================================
@dataclass
class Proof:
    LRO: List[Affine]
    Z: Affine
    H: List[Affine]
    batched_H: Affine
    claimed_values: List[int]      # [foldedH, linPol, l, r, o, s1, s2](zeta)
    zshift_H: Affine
    zshift_value: int

    def to_bytes(self) -> bytes:
        """Proof.WriteTo (gnark v0.8.0): 7 compressed G1 || BatchedProof.H || u32 len || 7 fr || ZShifted.H || fr."""
        out = b"".join(g1_compress(p) for p in self.LRO + [self.Z] + self.H)
        out += g1_compress(self.batched_H)
        out += len(self.claimed_values).to_bytes(4, "big")
        out += b"".join(o.fr_be_bytes(v) for v in self.claimed_values)
        out += g1_compress(self.zshift_H) + o.fr_be_bytes(self.zshift_value)
        return out

    @staticmethod
    def from_bytes(b: bytes) -> "Proof":
        pts = [g1_decompress(b[32 * i: 32 * i + 32]) for i in range(8)]
        k = int.from_bytes(b[256:260], "big")
        vals = [int.from_bytes(b[260 + 32 * i: 292 + 32 * i], "big") for i in range(k)]
        off = 260 + 32 * k
        return Proof(pts[0:3], pts[3], pts[4:7], pts[7], vals, g1_decompress(b[off:off + 32]),
                     int.from_bytes(b[off + 32: off + 64], "big"))


class BlindingStream:
    """Deterministic stand-in for crypto/rand.Reader feeding fr.SetRandom: 32 bytes, 4 LE limbs,
    top 2 bits cleared, rejection-sampled, limbs used as the MONTGOMERY representation unchanged."""

    def __init__(self, seed: int):
        self.state = seed & o.MASK64
        self.drawn: List[int] = []      # values (regular form) in draw order

    def next_mont(self) -> int:
        while True:
            v = 0
            for k in range(4):
                self.state, z = o.splitmix64(self.state)
                v |= z << (64 * k)
            v &= (1 << 254) - 1
            if v < R:
                return v

    def set_random(self) -> int:
        v = self.next_mont() * o.FR_RINV % R
        self.drawn.append(v)
        return v


def blind(poly: List[int], order: int, rng: BlindingStream) -> List[int]:
    """iop.Polynomial.Blind(order): p + (X^n - 1) * b(X), deg b = order."""
    n = len(poly)
    out = list(poly) + [0] * (order + 1)
    for i in range(order + 1):
        r = rng.set_random()
        out[i] = (out[i] - r) % R
        out[i + n] = (out[i + n] + r) % R
    return out


def solve(cs: SparseR1CS, witness: Sequence[int]) -> List[int]:
    """spr.Solve: every wire is supplied by ACVM, so this only checks the constraints."""
    sol = [w % R for w in witness]
    for k, g in enumerate(cs.gates):
        v = (g.ql * sol[g.a] + g.qr * sol[g.b] + g.qo * sol[g.c] + g.qm * sol[g.a] * sol[g.b] + g.qk) % R
        if v != 0:
            raise ValueError("constraint #%d is not satisfied" % k)
    return sol


def prove(cs: SparseR1CS, pk: ProvingKey, srs: SRS, full_witness: Sequence[int], rng: BlindingStream,
          trace: Optional[dict] = None) -> Proof:
    n, N4 = pk.n, pk.n_big
    dom = o.Domain(n)
    vk = pk.vk
    u = vk.coset_shift
    fs = Transcript("gamma", "beta", "alpha", "zeta")
    sol = solve(cs, full_witness)

    # evaluateLROSmallDomain
    s0 = sol[0] if sol else 0
    l = [s0] * n
    r_ = [s0] * n
    o_ = [s0] * n
    for i in range(cs.nb_public):
        l[i] = sol[i]
    off = cs.nb_public
    for i, g in enumerate(cs.gates):
        l[off + i], r_[off + i], o_[off + i] = sol[g.a], sol[g.b], sol[g.c]

    bl = blind(to_canonical(l), 1, rng)
    br = blind(to_canonical(r_), 1, rng)
    bo = blind(to_canonical(o_), 1, rng)
    LRO = [commit(bl, srs), commit(br, srs), commit(bo, srs)]

    # bindPublicData + gamma, beta
    for pt in vk.S + [vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk]:
        fs.bind("gamma", g1_marshal(pt))
    for i in range(cs.nb_public):
        fs.bind("gamma", o.fr_be_bytes(sol[i]))
    gamma = derive_randomness(fs, "gamma", *LRO)
    beta = challenge_to_fr(fs.compute("beta"))

    # BuildRatioCopyConstraint
    ident = identity_support(dom)
    wires = (l, r_, o_)
    num = [1] * n
    den = [1] * n
    for j in range(n):
        a = b = 1
        for k in range(3):
            a = a * (wires[k][j] + beta * ident[k * n + j] + gamma) % R
            b = b * (wires[k][j] + beta * ident[pk.permutation[k * n + j]] + gamma) % R
        num[j], den[j] = a, b
    z = [1] * n
    for j in range(n - 1):
        z[j + 1] = z[j] * num[j] % R * pow(den[j], -1, R) % R
    bz = blind(to_canonical(z), 2, rng)
    Z = commit(bz, srs)
    alpha = derive_randomness(fs, "alpha", Z)

    # qk completed with the public inputs
    qk_l = list(pk.lqk)
    for i in range(cs.nb_public):
        qk_l[i] = sol[i]
    qk_c = to_canonical(qk_l)

    # everything to Lagrange-coset on Domain[1] (bit-reversed layout)
    el, er, eo = (to_lagrange_coset_bitrev(x, N4) for x in (bl, br, bo))
    eqk = to_lagrange_coset_bitrev(qk_c, N4)
    eid = to_lagrange_coset_bitrev([0, 1], N4)
    ez = to_lagrange_coset_bitrev(bz, N4)
    lone = [0] * n
    lone[0] = 1
    elone = to_lagrange_coset_bitrev(to_canonical(lone), N4)
    log4 = N4.bit_length() - 1
    ratio = N4 // n
    uu = u * u % R

    # iop.Evaluate(fm, ...) + DivideByXMinusOne
    dom4 = o.Domain(N4)
    xn_inv = []
    un = pow(u, n, R)
    w4n_n = pow(dom4.generator, n, R)
    for i in range(ratio):
        xn_inv.append(pow((un * pow(w4n_n, i, R) - 1) % R, -1, R))
    t = [0] * N4
    for i in range(N4):
        nat = o.bit_reverse_index(i, log4)
        ishift = o.bit_reverse_index((nat + ratio) % N4, log4)    # z(wX): shift by one step of w_n
        L_, R_, O_ = el[i], er[i], eo[i]
        ic = (pk.l_ql[i] * L_ + pk.l_qr[i] * R_ + pk.l_qm[i] * L_ % R * R_ + pk.l_qo[i] * O_ + eqk[i]) % R
        fid = eid[i]
        a = (beta * fid + L_ + gamma) * (beta * u % R * fid + R_ + gamma) % R * (beta * uu % R * fid + O_ + gamma) % R * ez[i] % R
        b = (beta * pk.l_s1[i] + L_ + gamma) * (beta * pk.l_s2[i] + R_ + gamma) % R * (beta * pk.l_s3[i] + O_ + gamma) % R * ez[ishift] % R
        perm_term = (b - a) % R
        one_term = (ez[i] - 1) * elone[i] % R
        c = ((one_term * alpha + perm_term) % R * alpha + ic) % R
        t[i] = c * xn_inv[nat % ratio] % R
    h = _fft(t, True, o.DIT, True)      # bit-reversed Lagrange-coset -> canonical regular (FFTInverse DIT, coset)
    assert all(x == 0 for x in h[3 * (n + 2):]), "quotient degree too large (constraints not satisfied?)"
    h1, h2, h3 = h[: n + 2], h[n + 2: 2 * (n + 2)], h[2 * (n + 2): 3 * (n + 2)]
    H = [commit(h1, srs), commit(h2, srs), commit(h3, srs)]
    zeta = derive_randomness(fs, "zeta", *H)

    blzeta, brzeta, bozeta = eval_poly(bl, zeta), eval_poly(br, zeta), eval_poly(bo, zeta)
    zeta_shifted = zeta * vk.generator % R
    zu = eval_poly(bz, zeta_shifted)
    zshift_H = commit(divide_by_x_minus_a(bz, zu, zeta_shifted), srs)

    # computeLinearizedPolynomial
    rl = brzeta * blzeta % R
    s1z, s2z = eval_poly(pk.s1, zeta), eval_poly(pk.s2, zeta)
    c1 = (s1z * beta + blzeta + gamma) * (s2z * beta + brzeta + gamma) % R * zu % R * beta % R
    c2 = (beta * zeta + blzeta + gamma) * (beta * u % R * zeta + brzeta + gamma) % R * (beta * uu % R * zeta + bozeta + gamma) % R
    c2 = (-c2) % R
    lag = (pow(zeta, n, R) - 1) * pow((zeta - 1) % R, -1, R) % R * alpha % R * alpha % R * vk.size_inv % R
    lin = [0] * len(bz)
    for i in range(len(bz)):
        v = bz[i] * c2 % R
        if i < n:
            v = (v + pk.s3[i] * c1) % R
        v = v * alpha % R
        if i < n:
            v = (v + pk.qm[i] * rl + pk.ql[i] * blzeta + pk.qr[i] * brzeta + pk.qo[i] * bozeta + pk.cqk[i]) % R
        lin[i] = (v + bz[i] * lag) % R
    lin_digest = commit(lin, srs)

    zpm = pow(zeta, n + 2, R)
    folded_digest = o.g1_add(o.g1_mul(o.g1_add(o.g1_mul(H[2], zpm), H[1]), zpm), H[0])
    folded_h = [((h3[i] * zpm + h2[i]) % R * zpm + h1[i]) % R for i in range(n + 2)]

    polys = [folded_h, lin, bl, br, bo, pk.s1, pk.s2]
    digests = [folded_digest, lin_digest, LRO[0], LRO[1], LRO[2], vk.S[0], vk.S[1]]
    claimed = [eval_poly(p_, zeta) for p_ in polys]
    g_fs = Transcript("gamma")
    g_fs.bind("gamma", o.fr_be_bytes(zeta))
    for d in digests:
        g_fs.bind("gamma", g1_marshal(d))
    gk = challenge_to_fr(g_fs.compute("gamma"))
    folded_eval = 0
    for v in reversed(claimed):
        folded_eval = (folded_eval * gk + v) % R
    largest = max(len(p_) for p_ in polys)
    folded = list(polys[0]) + [0] * (largest - len(polys[0]))
    acc = gk
    for p_ in polys[1:]:
        for j, cj in enumerate(p_):
            folded[j] = (folded[j] + cj * acc) % R
        acc = acc * gk % R
    batched_H = commit(divide_by_x_minus_a(folded, folded_eval, zeta), srs)

    if trace is not None:
        trace.update(dict(gamma=gamma, beta=beta, alpha=alpha, zeta=zeta, bl=bl, br=br, bo=bo, bz=bz, z=z, h=h,
                          lin=lin, folded_h=folded_h, qk_c=qk_c, l=l, r=r_, o=o_, kzg_gamma=gk, t=t))
    return Proof(LRO, Z, H, batched_H, claimed, zshift_H, zu)


# ======================================================================================================
# 5. Verify (gnark v0.8.0 plonk.Verify) with an independent optimal-ate pairing
# ======================================================================================================
def verify(proof: Proof, vk: VerifyingKey, public_witness: Sequence[int], srs_g2) -> bool:
    fs = Transcript("gamma", "beta", "alpha", "zeta")
    for pt in vk.S + [vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk]:
        fs.bind("gamma", g1_marshal(pt))
    for w in public_witness:
        fs.bind("gamma", o.fr_be_bytes(w))
    gamma = derive_randomness(fs, "gamma", *proof.LRO)
    beta = challenge_to_fr(fs.compute("beta"))
    alpha = derive_randomness(fs, "alpha", proof.Z)
    zeta = derive_randomness(fs, "zeta", *proof.H)
    n = vk.size
    u = vk.coset_shift
    zeta_n = pow(zeta, n, R)
    zz = (zeta_n - 1) % R
    # PI(zeta) = sum_i L_i(zeta) w_i,  L_i(zeta) = w^i (zeta^n - 1) / (n (zeta - w^i))
    pi = 0
    for i, w in enumerate(public_witness):
        wi = pow(vk.generator, i, R)
        li = wi * zz % R * vk.size_inv % R * pow((zeta - wi) % R, -1, R) % R
        pi = (pi + li * w) % R
    lag1 = zz * vk.size_inv % R * pow((zeta - 1) % R, -1, R) % R
    q, lin_z, l, r, o_, s1, s2 = proof.claimed_values
    zu = proof.zshift_value
    t1 = (s1 * beta + l + gamma) * (s2 * beta + r + gamma) % R * (o_ + gamma) % R * alpha % R * zu % R
    lhs = (lin_z + pi + t1 - alpha * alpha % R * lag1) % R
    if lhs != q * zz % R:
        return False
    zpm = pow(zeta, n + 2, R)
    folded_h = o.g1_add(o.g1_mul(o.g1_add(o.g1_mul(proof.H[2], zpm), proof.H[1]), zpm), proof.H[0])
    uu = u * u % R
    c_s3 = (s1 * beta + l + gamma) * (s2 * beta + r + gamma) % R * beta % R * alpha % R * zu % R
    c_z = (beta * zeta + l + gamma) * (beta * u % R * zeta + r + gamma) % R * (beta * uu % R * zeta + o_ + gamma) % R
    c_z = (-c_z * alpha + alpha * alpha % R * lag1) % R
    lin_digest = None
    for pt, sc in ((vk.Ql, l), (vk.Qr, r), (vk.Qm, l * r % R), (vk.Qo, o_), (vk.Qk, 1), (vk.S[2], c_s3), (proof.Z, c_z)):
        lin_digest = o.g1_add(lin_digest, o.g1_mul(pt, sc))
    digests = [folded_h, lin_digest, proof.LRO[0], proof.LRO[1], proof.LRO[2], vk.S[0], vk.S[1]]
    g_fs = Transcript("gamma")
    g_fs.bind("gamma", o.fr_be_bytes(zeta))
    for d in digests:
        g_fs.bind("gamma", g1_marshal(d))
    gk = challenge_to_fr(g_fs.compute("gamma"))
    folded_digest = None
    folded_eval = 0
    acc = 1
    for d, v in zip(digests, proof.claimed_values):
        folded_digest = o.g1_add(folded_digest, o.g1_mul(d, acc))
        folded_eval = (folded_eval + v * acc) % R
        acc = acc * gk % R
    ok1 = kzg_verify(folded_digest, zeta, folded_eval, proof.batched_H, srs_g2)
    ok2 = kzg_verify(proof.Z, zeta * vk.generator % R, zu, proof.zshift_H, srs_g2)
    return ok1 and ok2


def kzg_verify(commitment: Affine, z: int, v: int, Hq: Affine, srs_g2) -> bool:
    """e(C - v G1 + z H, G2) == e(H, alpha G2)   <=>   e(C - vG1 + zH, G2) * e(-H, alpha G2) == 1."""
    lhs = o.g1_add(o.g1_add(commitment, o.g1_neg(o.g1_mul(o.G1_GEN, v))), o.g1_mul(Hq, z))
    return pairing_product_is_one([(lhs, srs_g2[0]), (o.g1_neg(Hq), srs_g2[1])])


# ------------------------------------------------------------------------------------------------------
# Fp2 / G2 / Fp12 / optimal ate pairing.  Fp12 = Fp[w] / (w^12 - 18 w^6 + 82), with u = w^6 - 9 (u^2 = -1).
# ------------------------------------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * d % P, (-a[1]) * d % P)


B2 = f2_mul((3, 0), f2_inv((9, 1)))       # twist curve y^2 = x^3 + 3/(9+u)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_mul((2, 0), y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(pt, k: int):
    acc = None
    base = pt
    k %= R
    while k:
        if k & 1:
            acc = g2_add(acc, base)
        base = g2_add(base, base)
        k >>= 1
    return acc


def g2_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):     # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


F12_ONE = [1] + [0] * 11


def f12_pow(a, e: int):
    res = F12_ONE
    base = a
    while e:
        if e & 1:
            res = f12_mul(res, base)
        base = f12_mul(base, base)
        e >>= 1
    return res


def _poly_deg(p):
    d = len(p) - 1
    while d >= 0 and p[d] == 0:
        d -= 1
    return d


def f12_inv(a):
    """extended Euclid in Fp[w] against the modulus polynomial"""
    mod = [82, 0, 0, 0, 0, 0, (-18) % P, 0, 0, 0, 0, 0, 1]
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], mod
    while _poly_deg(low) > 0:
        dl, dh = _poly_deg(low), _poly_deg(high)
        # r = high / low
        quo = [0] * 13
        temp = list(high)
        inv_lead = pow(low[dl], -1, P)
        for i in range(dh - dl, -1, -1):
            q = temp[dl + i] * inv_lead % P
            quo[i] = q
            for c in range(dl + 1):
                temp[c + i] = (temp[c + i] - low[c] * q) % P
        nm = list(hm)
        new = list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * quo[j]) % P
                new[i + j] = (new[i + j] - low[i] * quo[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, P)
    return [x * inv0 % P for x in lm[:12]]


def _f12_from_f2(c, shift: int):
    """(c0 + c1 u) * w^shift with u = w^6 - 9"""
    out = [0] * 12
    out[0] = (c[0] - 9 * c[1]) % P
    out[6] = c[1] % P
    if shift:
        wpow = [0] * 12
        wpow[shift] = 1
        out = f12_mul(out, wpow)
    return out


def _twist(q):
    # untwist: (x, y) on E'(Fp2) -> (x w^2, y w^3) on E(Fp12)
    return (_f12_from_f2(q[0], 2), _f12_from_f2(q[1], 3))


def _f12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def _f12_add(a, b): return [(x + y) % P for x, y in zip(a, b)]
def _f12_scalar(a, k): return [x * k % P for x in a]


def _line(p1, p2, t):
    """line through p1, p2 (points over Fp12) evaluated at t"""
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if x1 != x2:
        m = f12_mul(_f12_sub(y2, y1), f12_inv(_f12_sub(x2, x1)))
        return _f12_sub(f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    if y1 == y2:
        m = f12_mul(_f12_scalar(f12_mul(x1, x1), 3), f12_inv(_f12_scalar(y1, 2)))
        return _f12_sub(f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    return _f12_sub(xt, x1)


def _e12_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if y1 != y2:
            return None
        m = f12_mul(_f12_scalar(f12_mul(x1, x1), 3), f12_inv(_f12_scalar(y1, 2)))
    else:
        m = f12_mul(_f12_sub(y2, y1), f12_inv(_f12_sub(x2, x1)))
    x3 = _f12_sub(_f12_sub(f12_mul(m, m), x1), x2)
    y3 = _f12_sub(f12_mul(m, _f12_sub(x1, x3)), y1)
    return (x3, y3)


ATE_LOOP = 29793968203157093288      # 6x + 2


def miller_loop(q, p) -> List[int]:
    """f_{6x+2,Q}(P) with the two Frobenius line corrections (before the final exponentiation)."""
    if q is None or p is None:
        return F12_ONE
    Q = _twist(q)
    Pt = ([p[0]] + [0] * 11, [p[1]] + [0] * 11)
    Rp = Q
    f = F12_ONE
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        f = f12_mul(f12_mul(f, f), _line(Rp, Rp, Pt))
        Rp = _e12_add(Rp, Rp)
        if (ATE_LOOP >> i) & 1:
            f = f12_mul(f, _line(Rp, Q, Pt))
            Rp = _e12_add(Rp, Q)
    Q1 = (f12_pow(Q[0], P), f12_pow(Q[1], P))
    nQ2 = (f12_pow(Q1[0], P), [(-x) % P for x in f12_pow(Q1[1], P)])
    f = f12_mul(f, _line(Rp, Q1, Pt))
    Rp = _e12_add(Rp, Q1)
    f = f12_mul(f, _line(Rp, nQ2, Pt))
    return f


def final_exponentiation(f):
    return f12_pow(f, (P ** 12 - 1) // R)


def pairing(q, p):
    return final_exponentiation(miller_loop(q, p))


def pairing_product_is_one(pairs) -> bool:
    f = F12_ONE
    for p, q in pairs:
        f = f12_mul(f, miller_loop(q, p))
    return final_exponentiation(f) == F12_ONE


# ======================================================================================================
# 6. convenience: synthetic circuits and the reference's embedded fixtures
# ======================================================================================================
def synthetic_chain_circuit(nb_gates: int, seed: int, nb_public: int = 1):
    """x_{i+1} = x_i^2 + x_i + c_i   (qM=1, qL=1, qO=-1, qC=c_i), x_0 public (SURVEY.md §8d synthetic PLONK input)."""
    cs = SparseR1CS(nb_public=nb_public, nb_secret=nb_gates + 1 - nb_public)
    consts = o.random_fr(nb_gates, seed)
    x = [o.random_fr(1, seed ^ 0x5555)[0]]
    for i in range(nb_gates):
        x.append((x[i] * x[i] + x[i] + consts[i]) % R)
        cs.gates.append(Gate(1, 0, 1, R - 1, consts[i], i, i, i + 1))
    return cs, x


REFERENCE_FIXTURES = None  # filled lazily from /root/reference in tests/golden/make_golden.py, not at run time


# ======================================================================================================
# 7. the same prover through the C port (oracle/bn254_ref.c: oracle_plonk_prove) — CPU baseline + third opinion
# ======================================================================================================
class CProver:
    """Prepares the loaded-key arrays once (what gnark holds after ProvingKey.ReadFrom) and proves through C."""

    def __init__(self, cs: SparseR1CS, pk: ProvingKey, srs: SRS):
        import ctypes as C

        self.C = C
        self.lib = cref.load()
        self.cs, self.pk, self.srs = cs, pk, srs
        n, N4 = pk.n, pk.n_big
        dom = o.Domain(n)
        lone = to_lagrange_coset_bitrev([pow(n, -1, R)] * n, N4)     # L_1 = (1/n) sum X^i on the coset
        arr = lambda v: np.frombuffer(o.fr_to_mont_bytes(v), dtype=np.uint8).copy()
        self.polys = [arr(x) for x in (pk.ql, pk.qr, pk.qm, pk.qo, pk.cqk, pk.lqk, pk.s1, pk.s2, pk.s3)]
        self.cosets = [arr(x) for x in (pk.l_ql, pk.l_qr, pk.l_qm, pk.l_qo, pk.l_s1, pk.l_s2, pk.l_s3, lone)]
        self.perm = np.asarray(pk.permutation, dtype=np.int64)
        self.lro = np.asarray(pk.lro_wires, dtype=np.uint32)
        vk = pk.vk
        self.vk_points = np.frombuffer(o.g1_to_bytes(vk.S + [vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk]), dtype=np.uint8).copy()
        del dom

    @classmethod
    def from_arrays(cls, log2n: int, log2n_big: int, nb_public: int, nb_wires: int, polys, perm, lro, vk_points,
                    srs_g1_bytes, nthreads: int = 0) -> "CProver":
        """The same prover for sizes the Python setup cannot reach: `polys` are the 9 key polynomials as Montgomery
        byte images (Ql Qr Qm Qo CQk LQk S1 S2 S3); the Lagrange-coset forms are derived here with the C NTT, as
        gnark does when it loads a key."""
        import ctypes as C
        from types import SimpleNamespace

        self = object.__new__(cls)
        self.C, self.lib = C, cref.load()
        n, N4 = 1 << log2n, 1 << log2n_big
        self.polys = [np.frombuffer(bytes(p), dtype=np.uint8).copy() for p in polys]
        th = nthreads or cref.ncores()

        def coset(canonical: np.ndarray) -> np.ndarray:
            buf = np.zeros(N4 * 32, dtype=np.uint8)
            buf[: canonical.size] = canonical
            return np.frombuffer(cref.ntt(buf, log2n_big, 0, o.DIF, 1, th), dtype=np.uint8).copy()

        lone = np.tile(np.frombuffer(o.fr_to_mont_bytes([pow(n, -1, R)]), dtype=np.uint8), n)
        self.cosets = [coset(self.polys[i]) for i in (0, 1, 2, 3, 6, 7, 8)] + [coset(lone)]
        self.perm = np.ascontiguousarray(perm, dtype=np.int64)
        self.lro = np.ascontiguousarray(lro, dtype=np.uint32)
        self.vk_points = np.frombuffer(bytes(vk_points), dtype=np.uint8).copy()
        self.cs = SimpleNamespace(nb_public=nb_public, nb_secret=nb_wires - nb_public)
        self.pk = SimpleNamespace(n=n, n_big=N4)
        self.srs = SimpleNamespace(g1_bytes=np.ascontiguousarray(srs_g1_bytes))
        return self

    def prove_blob(self, full_witness, blinding_images: bytes, nthreads: int = 0) -> bytes:
        C = self.C
        sol = np.frombuffer(o.fr_to_mont_bytes(full_witness), dtype=np.uint8).copy() if not isinstance(full_witness, np.ndarray) else full_witness
        bl = np.frombuffer(bytes(blinding_images), dtype=np.uint8).copy()
        out = np.zeros(832, dtype=np.uint8)
        P = (C.c_void_p * 9)(*[a.ctypes.data for a in self.polys])
        Q = (C.c_void_p * 8)(*[a.ctypes.data for a in self.cosets])
        n = self.pk.n
        rc = self.lib.oracle_plonk_prove(n.bit_length() - 1, self.pk.n_big.bit_length() - 1, self.cs.nb_public,
                                         self.cs.nb_public + self.cs.nb_secret, P, Q, self.perm.ctypes.data,
                                         self.lro.ctypes.data, self.vk_points.ctypes.data, self.srs.g1_bytes.ctypes.data,
                                         sol.ctypes.data, bl.ctypes.data, nthreads or cref.ncores(), out.ctypes.data)
        assert rc == 0
        return out.tobytes()

    def prove(self, full_witness, blinding_images: bytes, nthreads: int = 0) -> Proof:
        return proof_from_blob(self.prove_blob(full_witness, blinding_images, nthreads))


def proof_from_blob(blob: bytes) -> Proof:
    """832-byte in-memory proof image (9 G1Affine + 8 fr.Element) -> Proof"""
    pts = o.g1_from_bytes(blob[:576])
    vals = o.fr_from_mont_bytes(blob[576:])
    return Proof(pts[0:3], pts[3], pts[4:7], pts[7], vals[:7], pts[8], vals[7])
