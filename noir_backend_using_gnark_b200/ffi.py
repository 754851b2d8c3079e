"""The reference's Rust-side PLONK wrapper (/root/reference/src/gnark_backend_wrapper/plonk/mod.rs:64-254) mirrored over
ctypes: same function names, arguments and payload encodings, calling the four cgo symbols that
lib/libgnark_backend_b200.so exports in place of the Go archive (include/gnark_backend_ffi.h).

Like the Go library, the C++ side reports every failure with a message on stderr and exit(1) (log.Fatal); callers that
need to survive a malformed payload run these functions in a child process, as the tests do."""
from __future__ import annotations

import ctypes as C
import json
import re
from pathlib import Path
from typing import Sequence, Tuple

from . import _lib

FFI_LIB_PATH = _lib.PKG / "lib" / "libgnark_backend_b200.so"
FFI_HEADER = _lib.PKG.parent / "include" / "gnark_backend_ffi.h"
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class GoString(C.Structure):
    """c_go_structures.rs:5-10: {ptr, length}, passed by value."""
    _fields_ = [("p", C.c_char_p), ("n", C.c_ssize_t)]

    @classmethod
    def of(cls, data: bytes) -> "GoString":
        s = cls(data, len(data))
        s._keep = data
        return s


class KeyPair(C.Structure):
    """c_go_structures.rs:22-26."""
    _fields_ = [("proving_key", C.c_void_p), ("verifying_key", C.c_void_p)]


_ffi = None


def ffi_header_symbols() -> list:
    text = re.sub(r"/\*.*?\*/", "", FFI_HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(Plonk[A-Za-z]+)\s*\(", text)))


def load_ffi() -> C.CDLL:
    global _ffi
    if _ffi is not None:
        return _ffi
    _lib.load()  # builds both libraries when they did not travel; no CPU fallback
    if not FFI_LIB_PATH.exists():
        raise ImportError("%s is missing: run `python -m noir_backend_using_gnark_b200.build`" % FFI_LIB_PATH)
    lib = C.CDLL(str(FFI_LIB_PATH))
    lib.PlonkProveWithPK.restype = C.c_void_p
    lib.PlonkProveWithPK.argtypes = [GoString, GoString, GoString]
    lib.PlonkVerifyWithMeta.restype = C.c_uint8
    lib.PlonkVerifyWithMeta.argtypes = [GoString, GoString, GoString]
    lib.PlonkVerifyWithVK.restype = C.c_uint8
    lib.PlonkVerifyWithVK.argtypes = [GoString, GoString, GoString, GoString]
    lib.PlonkPreprocess.restype = KeyPair
    lib.PlonkPreprocess.argtypes = [GoString, GoString]
    lib.b200zk_ffi_test_seed_blinding.restype = None
    lib.b200zk_ffi_test_seed_blinding.argtypes = [C.c_uint64, C.c_int]
    _ffi = lib
    return lib


def seed_blinding(seed=None) -> None:
    """TEST-ONLY: pin the prover's 9 blinding draws to a SplitMix64 stream (seed) so that proof bytes can be compared with
    the CPU checker's; None returns to /dev/urandom.  Seeded proofs are not zero-knowledge."""
    lib = load_ffi()
    lib.b200zk_ffi_test_seed_blinding(0 if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF, 0 if seed is None else 1)


def encode_felts(values: Sequence[int]) -> str:
    """serialize.rs:33-47 encode_felts: hex(u32-BE count || 32-byte big-endian field elements)."""
    return (len(values).to_bytes(4, "big") + b"".join((int(v) % FR_MODULUS).to_bytes(32, "big") for v in values)).hex()


def _cstr(ptr) -> str:
    return C.string_at(ptr).decode()  # the reference never frees these either (CStr::from_ptr, mod.rs:106)


def prove_with_pk(circuit_json: str, values: Sequence[int], proving_key: bytes) -> bytes:
    """mod.rs:64-99."""
    return bytes.fromhex(prove_with_pk_encoded(circuit_json.encode(), encode_felts(values).encode(), proving_key.hex().encode()))


def prove_with_pk_encoded(acir_json: bytes, values_hex: bytes, proving_key_hex: bytes) -> str:
    """The bare PlonkProveWithPK call on payloads that are already encoded (what benchmarks time)."""
    lib = load_ffi()
    return _cstr(lib.PlonkProveWithPK(GoString.of(acir_json), GoString.of(values_hex), GoString.of(proving_key_hex)))


def verify_with_vk_encoded(acir_json: bytes, proof_hex: bytes, public_inputs_hex: bytes, verifying_key_hex: bytes) -> int:
    """The bare PlonkVerifyWithVK call on encoded payloads."""
    lib = load_ffi()
    return int(lib.PlonkVerifyWithVK(GoString.of(acir_json), GoString.of(proof_hex), GoString.of(public_inputs_hex),
                                     GoString.of(verifying_key_hex)))


def preprocess_encoded(acir_json: bytes, quoted_values_hex: bytes) -> Tuple[str, str]:
    """The bare PlonkPreprocess call: (proving key hex, verifying key hex)."""
    lib = load_ffi()
    kp = lib.PlonkPreprocess(GoString.of(acir_json), GoString.of(quoted_values_hex))
    return _cstr(kp.proving_key), _cstr(kp.verifying_key)


def verify_with_meta(circuit_json: str, proof: bytes, public_inputs: Sequence[int]) -> bool:
    """mod.rs:101-141 (the Go side is a stub returning false, main.go:39-42)."""
    lib = load_ffi()
    r = lib.PlonkVerifyWithMeta(GoString.of(circuit_json.encode()), GoString.of(json.dumps(encode_felts(public_inputs)).encode()),
                                GoString.of(proof))
    return r == 1


def verify_with_vk(circuit_json: str, proof: bytes, public_inputs: Sequence[int], verifying_key: bytes) -> bool:
    """mod.rs:143-186."""
    r = verify_with_vk_encoded(circuit_json.encode(), proof.hex().encode(), encode_felts(public_inputs).encode(),
                               verifying_key.hex().encode())
    if r not in (0, 1):
        raise ValueError("VerifyInvalidBoolError")
    return r == 1


def get_exact_circuit_size(circuit_json: str) -> int:
    """mod.rs:188-193 over gnark_backend_wrapper/mod.rs:56-73 num_constraints: every opcode counts once, an arithmetic
    opcode additionally once per multiplication term plus once for its linear combination; opcodes other than
    Arithmetic / Directive are an error."""
    ops = json.loads(circuit_json)["opcodes"]
    total = len(ops)
    for op in ops:
        if "Arithmetic" in op:
            total += len(op["Arithmetic"]["mul_terms"]) + 1
        elif "Directive" not in op:
            raise ValueError("UnsupportedOpcodeError: %s" % next(iter(op)))
    return total


def preprocess(circuit_json: str, random_value: int = 1) -> Tuple[bytes, bytes]:
    """mod.rs:195-240: (proving_key, verifying_key) bytes.  The Rust side sends num_witnesses - 1 copies of ONE random
    field element (`vec![rand::random(); n]`), JSON-quoted."""
    m = re.search(r'"current_witness_index"\s*:\s*(\d+)', circuit_json)
    n = int(m.group(1)) if m else int(json.loads(circuit_json)["current_witness_index"])
    pk, vk = preprocess_encoded(circuit_json.encode(), json.dumps(encode_felts([random_value] * n)).encode())
    return bytes.fromhex(pk), bytes.fromhex(vk)
