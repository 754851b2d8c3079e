"""CPU tests of the outer boundary (include/gnark_backend_ffi.h, lib/libgnark_backend_b200.so): the library exports the
four cgo symbols of /root/reference/gnark_backend_ffi/main.go, and PlonkVerifyWithVK — CPU work in the reference too —
accepts / rejects proofs made by the oracle prover when fed the Rust crate's string payloads.  Proving and preprocessing
need the device and are covered by tests/test_ffi_gpu.py."""
import ctypes as C

import pytest

from noir_backend_using_gnark_b200 import ffi
from oracle import bn254 as o
from oracle import ffi_formats as ff
from oracle import plonk as pl

from .ffi_util import run_child, write_srs_file
from .test_plonk_oracle import FIXTURES


def test_exports_the_cgo_symbols(lib_built):
    names = ffi.ffi_header_symbols()
    assert names == ["PlonkPreprocess", "PlonkProveWithPK", "PlonkVerifyWithMeta", "PlonkVerifyWithVK"]  # main.go:24,39,44,58
    lib = C.CDLL(str(ffi.FFI_LIB_PATH))
    for n in names:
        assert hasattr(lib, n), n


def test_felt_encoding_matches_the_rust_serializer():
    # serialize.rs:33-47: u32-BE count then 32-byte big-endian elements
    assert ffi.encode_felts([]) == "00000000"
    assert ffi.encode_felts([1, o.R_MOD - 1]) == "00000002" + "00" * 31 + "01" + (o.R_MOD - 1).to_bytes(32, "big").hex()
    assert ffi.encode_felts([5, 6]) == ff.felts_hex([5, 6])


def test_exact_circuit_size_counts_like_the_wrapper():
    # gnark_backend_wrapper/mod.rs:56-73: opcodes + (mul terms + 1) per arithmetic opcode
    js = FIXTURES[0][0]
    assert ffi.get_exact_circuit_size(js) == 5 + (0 + 1) + (1 + 1) + (1 + 1) + (0 + 1)
    assert ffi.get_exact_circuit_size('{"current_witness_index":0,"opcodes":[],"public_inputs":[]}') == 0   # plonk/mod.rs:249-253
    with pytest.raises(ValueError):
        ffi.get_exact_circuit_size('{"current_witness_index":0,"opcodes":[{"BlackBoxFuncCall":{}}],"public_inputs":[]}')


@pytest.fixture(scope="module")
def proved(tmp_path_factory, lib_built):
    home = tmp_path_factory.mktemp("cfg")
    srs = pl.SRS(128, 0xB2000005)
    write_srs_file(home, srs)
    cases = []
    for js, vals in FIXTURES:
        vals = [v % o.R_MOD for v in vals]
        cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
        pk = pl.setup(cs, srs)
        proof = pl.prove(cs, pk, srs, pub + sec, pl.BlindingStream(0xB2000006))
        cases.append((js, vals, ff.vk_bytes(pk.vk), proof.to_bytes()))
    return home, cases


def test_verify_with_vk_accepts_oracle_proofs_and_rejects_tampering(proved):
    home, cases = proved
    steps = []
    for js, vals, vk, proof in cases:
        strv = [str(v) for v in vals]
        steps.append({"op": "verify", "acir": js, "values": strv, "vk": vk.hex(), "proof": proof.hex()})
        bad = bytearray(proof)
        bad[300] ^= 1                              # a claimed value
        steps.append({"op": "verify", "acir": js, "values": strv, "vk": vk.hex(), "proof": bytes(bad).hex()})
        badv = list(vals)
        badv[1] = (badv[1] + 1) % o.R_MOD          # witness 2: the public input of the first two circuits
        steps.append({"op": "verify", "acir": js, "values": [str(v) for v in badv], "vk": vk.hex(), "proof": proof.hex()})
        steps.append({"op": "verify_meta", "acir": js, "values": strv, "proof": proof.hex()})
    rc, res, err = run_child(steps, home)
    assert rc == 0, err
    # third circuit has no public inputs, so changing witness 2 does not change the statement
    assert res == [True, False, False, False, True, False, False, False, True, False, True, False]


@pytest.mark.parametrize("field,value,needle", [
    ("acir", "{not json", "ACIR"),
    ("proof", "zz", "hex"),
    ("proof", "00" * 40, "proof"),
    ("vk", "00" * 10, "verifying key"),
])
def test_malformed_payloads_are_fatal_like_log_fatal(proved, field, value, needle):
    home, cases = proved
    js, vals, vk, proof = cases[1]
    step = {"op": "raw_verify", "acir": js, "values": ff.felts_hex(vals), "vk": vk.hex(), "proof": proof.hex()}
    rc, res, err = run_child([step], home)
    assert rc == 0 and res == [1], err
    step[field] = value
    rc, res, err = run_child([step], home)
    assert rc == 1 and needle in err, (rc, err)          # main.go:29,49,64,71: log.Fatal -> exit status 1


def test_mutated_payloads_never_crash(proved):
    """Byte-level mutations of every PlonkVerifyWithVK payload: the call must end as the reference would — a 0/1 answer
    or log.Fatal's exit status 1 — never by a signal."""
    import random

    home, cases = proved
    js, vals, vk, proof = cases[0]
    base = {"op": "raw_verify", "acir": js, "values": ff.felts_hex(vals), "vk": vk.hex(), "proof": proof.hex()}
    rng = random.Random(0xB200F)
    outcomes = set()
    for trial in range(14):
        step = dict(base)
        field = ["acir", "values", "vk", "proof"][trial % 4]
        s = step[field]
        kind = rng.randrange(4)
        pos = rng.randrange(len(s))
        if kind == 0:
            s = s[:pos]                                              # truncation
        elif kind == 1:
            s = s[:pos] + rng.choice("0123456789abcdef{}[]\",:x") + s[pos + 1:]   # one character replaced
        elif kind == 2:
            s = s[:pos] + s[pos:pos + 7] * 3 + s[pos:]               # a run duplicated
        else:
            s = s + s[:pos]                                          # trailing garbage
        step[field] = s
        rc, res, err = run_child([step], home)
        assert rc in (0, 1), (field, kind, rc, err[-200:])
        if rc == 0:
            assert res in ([0], [1])
        outcomes.add((rc, tuple(res) if res else None))
    assert len(outcomes) >= 2
