"""GPU tests of the outer boundary: the reference's three embedded circuits (main.go:233-247) through the string FFI the
Rust crate uses — PlonkPreprocess -> PlonkProveWithPK -> PlonkVerifyWithVK — with the SRS cache file the reference
reads, compared byte for byte with the oracle's keys and proof (B200ZK_BLINDING_SEED pins the 9 blinding draws)."""
import os

import pytest

from oracle import bn254 as o
from oracle import ffi_formats as ff
from oracle import plonk as pl

from .ffi_util import run_child, write_srs_file
from .test_plonk_oracle import FIXTURES

pytestmark = pytest.mark.gpu
SEED = 0xB2000006


@pytest.fixture(scope="module")
def home(tmp_path_factory, lib_built):
    d = tmp_path_factory.mktemp("cfg")
    write_srs_file(d, pl.SRS(128, 0xB2000005))
    return d


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_preprocess_prove_verify_match_the_oracle_bytes(home, idx):
    js, vals = FIXTURES[idx]
    vals = [v % o.R_MOD for v in vals]
    srs = pl.SRS(128, 0xB2000005)
    cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
    pk = pl.setup(cs, srs)
    want_proof = pl.prove(cs, pk, srs, pub + sec, pl.BlindingStream(SEED)).to_bytes()
    strv = [str(v) for v in vals]
    # one process: keys, proof, verification (the Rust backend's lifecycle)
    rc, res, err = run_child([{"op": "preprocess", "acir": js, "random_value": 12345}], home)
    assert rc == 0, err
    keys = res[0]
    assert bytes.fromhex(keys["vk"]) == ff.vk_bytes(pk.vk)
    assert bytes.fromhex(keys["pk"]) == ff.pk_bytes(pk)
    # separate processes for prove and verify: nothing but the strings and the SRS file carries over
    rc, res, err = run_child([{"op": "prove", "acir": js, "values": strv, "pk": keys["pk"]}], home,
                             {"B200ZK_BLINDING_SEED": str(SEED)})
    assert rc == 0, err
    assert bytes.fromhex(res[0]) == want_proof
    rc, res, err = run_child([{"op": "verify", "acir": js, "values": strv, "vk": keys["vk"], "proof": res[0]}], home)
    assert rc == 0 and res == [True], err


def test_unseeded_proofs_differ_and_verify(home):
    js, vals = FIXTURES[1]
    strv = [str(v % o.R_MOD) for v in vals]
    rc, res, err = run_child([
        {"op": "preprocess", "acir": js},
        {"op": "prove", "acir": js, "values": strv, "pk": "@0.pk"},
        {"op": "prove", "acir": js, "values": strv, "pk": "@0.pk"},
    ], home)
    assert rc == 0, err
    p1, p2 = res[1], res[2]
    assert p1 != p2 and len(bytes.fromhex(p1)) == 548          # crypto/rand blinding
    steps = [{"op": "verify", "acir": js, "values": strv, "vk": res[0]["vk"], "proof": p} for p in (p1, p2)]
    rc, res, err = run_child(steps, home)
    assert rc == 0 and res == [True, True], err


def test_unsatisfied_witness_is_fatal(home):
    js, vals = FIXTURES[1]
    bad = [v % o.R_MOD for v in vals]
    bad[0] = 7                                                  # 7 != 2
    steps = [{"op": "preprocess", "acir": js}, {"op": "prove", "acir": js, "values": [str(v) for v in bad], "pk": "@0.pk"}]
    rc, res, err = run_child(steps, home)
    assert rc == 1 and "constraint #0" in err, (rc, err)        # plonk.go:67-70: plonk.Prove error -> log.Fatal


def test_foreign_or_truncated_proving_key_is_fatal(home):
    js, vals = FIXTURES[1]
    strv = [str(v % o.R_MOD) for v in vals]
    rc, res, err = run_child([{"op": "preprocess", "acir": js}], home)
    assert rc == 0, err
    pk = res[0]["pk"]
    rc, _, err = run_child([{"op": "prove", "acir": js, "values": strv, "pk": pk[:-64]}], home)
    assert rc == 1 and "EOF" in err, (rc, err)                  # helpers.go:55-58: pk.ReadFrom error -> log.Fatal
    other = pk[:200] + ("1" if pk[200] != "1" else "2") + pk[201:]   # a different commitment inside the embedded vk
    rc, _, err = run_child([{"op": "prove", "acir": js, "values": strv, "pk": other}], home)
    assert rc == 1 and "does not belong" in err, (rc, err)
    rc, _, err = run_child([{"op": "prove", "acir": js, "values": strv, "pk": pk[:-1] + "x"}], home)
    assert rc == 0 or "hex" in err                              # the tail of the payload is not re-read


def test_srs_is_generated_and_cached_when_missing(tmp_path, lib_built):
    js, vals = FIXTURES[0]
    strv = [str(v % o.R_MOD) for v in vals]
    env = {"B200ZK_SRS_SIZE": "256"}
    rc, res, err = run_child([
        {"op": "preprocess", "acir": js},
        {"op": "prove", "acir": js, "values": strv, "pk": "@0.pk"},
    ], tmp_path, env)
    assert rc == 0, err
    path = os.path.join(str(tmp_path), "noir-lang", "srs.hex")
    raw = bytes.fromhex(open(path).read())                      # common.go:107-125
    assert int.from_bytes(raw[:4], "big") == 256 and len(raw) == 4 + 32 * 256 + 128
    g1_0, g1_1 = pl.g1_decompress(raw[4:36]), pl.g1_decompress(raw[36:68])
    assert g1_0 == o.G1_GEN
    assert raw[4 + 32 * 256: 4 + 32 * 256 + 64] == ff.g2_compress(pl.G2_GEN)
    # a second process loads the file instead of drawing a new alpha: the proof of the first verifies against it
    rc, res2, err = run_child([{"op": "verify", "acir": js, "values": strv, "vk": res[0]["vk"], "proof": res[1]}], tmp_path, env)
    assert rc == 0 and res2 == [True], err
    assert bytes.fromhex(open(path).read()) == raw
    # the file is a consistent SRS: e(alpha G1, G2) == e(G1, alpha G2), alpha G2 read back through the oracle's own decoder
    x1 = int.from_bytes(bytes([raw[-64] & 0x3F]) + raw[-63:-32], "big")
    x0 = int.from_bytes(raw[-32:], "big")
    cands = [q for q in _g2_with_x((x0, x1))]
    assert any(pl.pairing_product_is_one([(g1_1, pl.G2_GEN), (o.g1_neg(g1_0), q)]) for q in cands)


def _g2_with_x(x):
    """Both points of the twist with abscissa x (test-side square root in Fp2 by brute exponentiation)."""
    P = o.P_MOD
    b = pl.f2_mul((3, 0), pl.f2_inv((9, 1)))
    rhs = pl.f2_add(pl.f2_mul(pl.f2_mul(x, x), x), b)
    # sqrt in Fp2 via the norm: standard complex method
    a0, a1 = rhs
    norm = (a0 * a0 + a1 * a1) % P
    s = pow(norm, (P + 1) // 4, P)
    assert s * s % P == norm
    for sign in (s, P - s):
        t = (a0 + sign) * pow(2, P - 2, P) % P
        r0 = pow(t, (P + 1) // 4, P)
        if r0 * r0 % P != t:
            continue
        r1 = a1 * pow(2 * r0, P - 2, P) % P
        y = (r0, r1)
        if pl.f2_mul(y, y) == rhs:
            return [(x, y), (x, pl.f2_sub((0, 0), y))]
    raise AssertionError("x is not on the twist")


def test_key_too_large_for_the_srs_is_fatal(home):
    # 128-point SRS, circuit of 200 rows -> kzg.Commit would fail with ErrInvalidPolynomialSize in plonk.Setup (plonk.go:21-24)
    ops = ",".join('{"Arithmetic":{"mul_terms":[],"linear_combinations":[["%s",1]],"q_c":"%s"}}' % ("0" * 63 + "1", "0" * 64)
                   for _ in range(200))
    js = '{"current_witness_index":1,"opcodes":[%s],"public_inputs":[]}' % ops
    rc, res, err = run_child([{"op": "preprocess", "acir": js}], home)
    assert rc == 1 and "SRS" in err, (rc, err)


def test_value_payload_edge_cases_on_the_device_decoder(home):
    """PlonkProveWithPK decodes the hex felts on the device: upper-case digits, values >= r (SetBytes reduces), bytes
    after the vector, an invalid character inside the vector (fatal like hex.DecodeString), a short vector
    (UnmarshalBinary's ignored error -> no values -> the constraints cannot hold)."""
    js, vals = FIXTURES[1]
    vals = [v % o.R_MOD for v in vals]
    srs = pl.SRS(128, 0xB2000005)
    cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
    pk = pl.setup(cs, srs)
    want = pl.prove(cs, pk, srs, pub + sec, pl.BlindingStream(SEED)).to_bytes().hex()
    env = {"B200ZK_BLINDING_SEED": str(SEED)}
    plain = ff.felts_hex(vals)
    unreduced = "%08x" % len(vals) + "".join("%064x" % (v + o.R_MOD) for v in vals)      # < 2^256, same residues
    steps = [{"op": "preprocess", "acir": js},
             {"op": "raw_prove", "acir": js, "values": plain.upper(), "pk": "@0.pk"},
             {"op": "raw_prove", "acir": js, "values": unreduced, "pk": "@0.pk"},
             {"op": "raw_prove", "acir": js, "values": plain + "00ff", "pk": "@0.pk"}]
    rc, res, err = run_child(steps, home, env)
    assert rc == 0, err
    assert res[1] == want and res[2] == want and res[3] == want
    bad = plain[:100] + "g" + plain[101:]
    rc, res, err = run_child(steps[:1] + [{"op": "raw_prove", "acir": js, "values": bad, "pk": "@0.pk"}], home, env)
    assert rc == 1 and "hex" in err, (rc, err)
    rc, res, err = run_child(steps[:1] + [{"op": "raw_prove", "acir": js, "values": plain + "zz", "pk": "@0.pk"}], home, env)
    assert rc == 1 and "hex" in err, (rc, err)
    rc, res, err = run_child(steps[:1] + [{"op": "raw_prove", "acir": js, "values": plain[:-64], "pk": "@0.pk"}], home, env)
    assert rc == 1, (rc, err)      # no values: different key size / unsatisfied system, fatal either way


def test_key_cache_eviction_and_alternating_circuits(home):
    """two circuits through one process with room for a single cached key (B200ZK_FFI_CACHE=1): every switch evicts the
    other circuit's parsed ACIR and device-resident key and rebuilds them; proofs stay byte-identical to the oracle's."""
    env = {"B200ZK_BLINDING_SEED": str(SEED), "B200ZK_FFI_CACHE": "1"}
    srs = pl.SRS(128, 0xB2000005)
    want, steps = {}, []
    for idx in (0, 2):
        js, vals = FIXTURES[idx]
        vals = [v % o.R_MOD for v in vals]
        cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
        pk = pl.setup(cs, srs)
        want[idx] = (pl.prove(cs, pk, srs, pub + sec, pl.BlindingStream(SEED)).to_bytes().hex(), ff.pk_bytes(pk).hex())
    order = [0, 2, 0, 2, 2, 0]
    for idx in order:
        js, vals = FIXTURES[idx]
        steps.append({"op": "prove", "acir": js, "values": [str(v % o.R_MOD) for v in vals], "pk": want[idx][1]})
    rc, res, err = run_child(steps, home, env)
    assert rc == 0, err
    assert res == [want[idx][0] for idx in order]
