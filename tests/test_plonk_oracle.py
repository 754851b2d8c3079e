"""CPU tests of the PLONK oracle (oracle/plonk.py) and of the product's host glue (package plonk.py):
the three embedded ACIR circuits of /root/reference/gnark_backend_ffi/main.go:233-247 must decode, build, prove and
verify; tampered proofs must be rejected; the product's vectorised permutation / row builder must agree with the
oracle's sequential restatement of gnark's buildPermutation."""
import numpy as np
import pytest

from noir_backend_using_gnark_b200 import plonk as zkp
from oracle import bn254 as o
from oracle import plonk as pl

M1 = "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000000"
ONE = "0000000000000000000000000000000000000000000000000000000000000001"
ZERO = "0" * 64


def _fixture(last_lin, last_qc, public_inputs):
    return (
        '{"current_witness_index":6,"opcodes":[{"Arithmetic":{"mul_terms":[],"linear_combinations":[["%s",1],["%s",2],["%s",3]],"q_c":"%s"}},'
        '{"Directive":{"Invert":{"x":3,"result":4}}},'
        '{"Arithmetic":{"mul_terms":[["%s",3,4]],"linear_combinations":[["%s",5]],"q_c":"%s"}},'
        '{"Arithmetic":{"mul_terms":[["%s",3,5]],"linear_combinations":[["%s",3]],"q_c":"%s"}},'
        '{"Arithmetic":{"mul_terms":[],"linear_combinations":[["%s",5]],"q_c":"%s"}}],"public_inputs":%s}'
        % (ONE, M1, M1, ZERO, ONE, M1, ZERO, ONE, M1, ZERO, last_lin, last_qc, public_inputs)
    )


# the three circuits + witness vectors embedded in the reference's main() (main.go:233-247)
FIXTURES = [
    (_fixture(M1, ONE, "[2]"), [0, 1, -1, -1, 1, 0]),     # 0 != 1
    (_fixture(ONE, ZERO, "[2]"), [2, 2, 0, 0, 0, 0]),     # 2 == 2
    (_fixture(ONE, ZERO, "[]"), [3, 3, 0, 0, 0, 0]),      # 3 == 3, no public input
]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_reference_fixture_proves_and_verifies(idx):
    js, vals = FIXTURES[idx]
    vals = [v % o.R_MOD for v in vals]
    cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
    assert len(cs.gates) == 4
    srs = pl.SRS(128, 0xB2000005)          # main.go:176 uses a 128-point SRS for these circuits
    pk = pl.setup(cs, srs)
    proof = pl.prove(cs, pk, srs, pub + sec, pl.BlindingStream(0xB2000006))
    assert len(proof.to_bytes()) == 548
    assert pl.Proof.from_bytes(proof.to_bytes()) == proof
    assert pl.verify(proof, pk.vk, pub, srs.g2)
    if pub:
        assert not pl.verify(proof, pk.vk, [(pub[0] + 1) % o.R_MOD], srs.g2)


def test_tampered_proof_rejected():
    cs, x = pl.synthetic_chain_circuit(20, 0xB2000004)
    srs = pl.SRS(64, 12345)
    pk = pl.setup(cs, srs)
    proof = pl.prove(cs, pk, srs, x, pl.BlindingStream(7))
    assert pl.verify(proof, pk.vk, x[:1], srs.g2)
    bad = pl.Proof.from_bytes(proof.to_bytes())
    bad.claimed_values[3] = (bad.claimed_values[3] + 1) % o.R_MOD
    assert not pl.verify(bad, pk.vk, x[:1], srs.g2)
    bad = pl.Proof.from_bytes(proof.to_bytes())
    bad.Z = o.g1_add(bad.Z, o.G1_GEN)
    assert not pl.verify(bad, pk.vk, x[:1], srs.g2)


def test_unsatisfied_witness_rejected_by_solver():
    cs, x = pl.synthetic_chain_circuit(5, 1)
    x[3] = (x[3] + 1) % o.R_MOD
    srs = pl.SRS(32, 5)
    pk = pl.setup(cs, srs)
    with pytest.raises(ValueError):
        pl.prove(cs, pk, srs, x, pl.BlindingStream(1))


def test_pairing_bilinearity():
    a, b = 0x1234567, 0x89ABCDEF01
    lhs = pl.pairing(pl.g2_mul(pl.G2_GEN, b), o.g1_mul(o.G1_GEN, a))
    rhs = pl.f12_pow(pl.pairing(pl.G2_GEN, o.G1_GEN), a * b % o.R_MOD)
    assert lhs == rhs and lhs != pl.F12_ONE
    assert pl.g2_is_on_curve(pl.G2_GEN) and pl.g2_mul(pl.G2_GEN, o.R_MOD - 1) == (pl.G2_GEN[0], pl.f2_sub((0, 0), pl.G2_GEN[1]))


def test_product_glue_matches_oracle_glue():
    for js, vals in FIXTURES:
        vals = [v % o.R_MOD for v in vals]
        cs_o, pub_o, sec_o = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
        cs_p, pub_p, sec_p = zkp.build_sparse_r1cs(js, vals)
        assert (pub_o, sec_o) == (pub_p, sec_p)
        assert (cs_o.nb_public, cs_o.nb_secret) == (cs_p.nb_public, cs_p.nb_secret)
        for k, g in enumerate(cs_o.gates):
            assert (g.ql, g.qr, g.qm, g.qo, g.qk, g.a, g.b, g.c) == (
                cs_p.ql[k], cs_p.qr[k], cs_p.qm[k], cs_p.qo[k], cs_p.qk[k], cs_p.a[k], cs_p.b[k], cs_p.c[k])


def test_vectorised_permutation_matches_gnark_cycle_walk():
    rng = np.random.default_rng(5)
    for n, wires in ((8, 3), (64, 10), (1024, 700)):
        lro = rng.integers(0, wires, size=3 * n).astype(np.uint32)
        got = zkp.build_permutation(lro)
        cycle = [-1] * wires
        perm = [-1] * (3 * n)
        for i in range(3 * n):
            if cycle[lro[i]] != -1:
                perm[i] = cycle[lro[i]]
            cycle[lro[i]] = i
        for i in range(3 * n):
            if perm[i] == -1:
                perm[i] = cycle[lro[i]]
        assert got.tolist() == perm


@pytest.mark.parametrize("gates,nb_public", [(1, 1), (13, 1), (300, 3)])
def test_c_prover_matches_python_prover(gates, nb_public):
    """third implementation: the C port (CPU baseline arm) must emit the same proof bytes as the Python oracle"""
    cs, x = pl.synthetic_chain_circuit(gates, 0xB2000004 + gates, nb_public)
    size = 1
    while size < gates + nb_public:
        size <<= 1
    srs = pl.SRS(size + 3, 0xB2000005)
    pk = pl.setup(cs, srs)
    want = pl.prove(cs, pk, srs, x, pl.BlindingStream(11)).to_bytes()
    st = pl.BlindingStream(11)
    blind = b"".join(o.limbs_le(st.next_mont()) for _ in range(9))
    got = pl.CProver(cs, pk, srs).prove(x, blind, nthreads=3)
    assert got.to_bytes() == want
    assert pl.verify(got, pk.vk, x[:nb_public], srs.g2)
