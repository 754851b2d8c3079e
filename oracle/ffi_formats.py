"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the string payloads of the reference's Rust<->Go FFI, restated in
Python so that tests can feed lib/libgnark_backend_b200.so exactly what the Rust crate sends and compare what comes back.

  felts          /root/reference/src/gnark_backend_wrapper/serialize.rs:33-47, gnark_backend_ffi/internal/backend/helpers.go:14-36
  keys, proof    helpers.go:38-94 (hex of gnark's WriteTo), gnark v0.8.0 layouts as recalled in SURVEY.md Appendix C
  SRS cache      gnark_backend_ffi/backend/common.go:78-144 (hex of kzg.SRS.WriteTo)

Parity unpinned against real gnark bytes (no Go toolchain, no golden vectors in the reference) — these layouts are the
same recollection the product follows; what the tests pin is that two independent implementations of it agree."""
from __future__ import annotations

import json
from typing import List, Sequence

from . import bn254 as o
from . import plonk as pl

P, R = o.P_MOD, o.R_MOD


def felts_hex(values: Sequence[int]) -> str:
    """serialize_felts: u32-BE count || 32-byte big-endian elements, hex-encoded."""
    return (len(values).to_bytes(4, "big") + b"".join((v % R).to_bytes(32, "big") for v in values)).hex()


def felts_hex_quoted(values: Sequence[int]) -> str:
    """What PlonkPreprocess receives: serde_json::to_string of the hex string (plonk/mod.rs:197-203)."""
    return json.dumps(felts_hex(values))


def _f2_lex_largest(y) -> bool:
    return y[1] > (P - 1) // 2 if y[1] != 0 else y[0] > (P - 1) // 2


def g2_compress(pt) -> bytes:
    """G2Affine.Bytes(): X.A1 || X.A0 big-endian, flags in the top two bits of the first byte."""
    if pt is None:
        return b"\x40" + b"\0" * 63
    (x0, x1), y = pt
    b = bytearray(x1.to_bytes(32, "big") + x0.to_bytes(32, "big"))
    b[0] |= 0xC0 if _f2_lex_largest(y) else 0x80
    return bytes(b)


def srs_bytes(srs: pl.SRS) -> bytes:
    """kzg.SRS.WriteTo: u32 len || compressed G1 powers || G2[0] || G2[1]."""
    pts = o.g1_from_bytes(srs.g1_bytes)
    return len(pts).to_bytes(4, "big") + b"".join(pl.g1_compress(p) for p in pts) + g2_compress(srs.g2[0]) + g2_compress(srs.g2[1])


def vk_bytes(vk: pl.VerifyingKey) -> bytes:
    """plonk VerifyingKey.WriteTo: Size u64 | SizeInv | Generator | NbPublicVariables u64 | S[0..2] | Ql Qr Qm Qo Qk."""
    out = vk.size.to_bytes(8, "big") + o.fr_be_bytes(vk.size_inv) + o.fr_be_bytes(vk.generator) + vk.nb_public.to_bytes(8, "big")
    return out + b"".join(pl.g1_compress(p) for p in list(vk.S) + [vk.Ql, vk.Qr, vk.Qm, vk.Qo, vk.Qk])


def _domain_bytes(n: int) -> bytes:
    """fft.Domain.WriteTo: Cardinality u64 | CardinalityInv | Generator | GeneratorInv | FrMultiplicativeGen | ...Inv."""
    d = o.Domain(n)
    return (d.cardinality.to_bytes(8, "big") + o.fr_be_bytes(d.cardinality_inv) + o.fr_be_bytes(d.generator)
            + o.fr_be_bytes(d.generator_inv) + o.fr_be_bytes(d.fr_multiplicative_gen) + o.fr_be_bytes(d.fr_multiplicative_gen_inv))


def _vector_bytes(v: Sequence[int]) -> bytes:
    return len(v).to_bytes(4, "big") + b"".join(o.fr_be_bytes(x) for x in v)


def pk_bytes(pk: pl.ProvingKey) -> bytes:
    """plonk ProvingKey.WriteTo: Vk | Domain[0] | Domain[1] | Ql Qr Qm Qo CQk LQk S1 S2 S3 | Permutation.
    The polynomials are []fr.Element (u32 count + elements); Permutation is a []int64 that gnark-crypto's Encoder passes
    to binary.Write unchanged: 3n big-endian int64, no count (ReadFrom sizes it from Domain[0].Cardinality)."""
    out = vk_bytes(pk.vk) + _domain_bytes(pk.n) + _domain_bytes(pk.n_big)
    for poly in (pk.ql, pk.qr, pk.qm, pk.qo, pk.cqk, pk.lqk, pk.s1, pk.s2, pk.s3):
        out += _vector_bytes(poly)
    out += b"".join(int(x).to_bytes(8, "big") for x in pk.permutation)
    return out
