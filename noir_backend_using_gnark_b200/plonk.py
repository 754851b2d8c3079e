"""Host side of the device-resident PLONK prover (C ABI: b200zk_plonk_* in include/b200zk.h).

Mirrors the reference's flow for this path:
  plonk_backend.BuildSparseR1CS  (/root/reference/gnark_backend_ffi/backend/plonk/sparse_r1cs.go:18-25, 44-107)
  plonk.Setup                    (call site backend/plonk/plonk.go:21)   -> ProvingKey.Setup
  plonk.Prove                    (call site backend/plonk/plonk.go:67)   -> ProvingKey.Prove
  backend_helpers.SerializeProof (internal/backend/helpers.go:75-80)     -> Proof.to_gnark_bytes
What stays on the CPU here is what the reference's Go glue does on the CPU: laying out the constraint rows and the
wire permutation (O(n)); every field / curve operation on polynomials runs in libb200zk.so.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .api import SRS, Context, default_context

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_MONT = 1 << 256
_R_INV_R = pow(_MONT, -1, R_MOD)
_R_INV_P = pow(_MONT, -1, P_MOD)


def fr_to_mont(vals: Sequence[int]) -> np.ndarray:
    """python ints -> (len*32,) uint8 in gnark's in-memory layout (Montgomery, 4 LE limbs)"""
    return np.frombuffer(b"".join(((v % R_MOD) * _MONT % R_MOD).to_bytes(32, "little") for v in vals), dtype=np.uint8)


def fr_from_mont(buf) -> List[int]:
    b = bytes(buf)
    return [int.from_bytes(b[i:i + 32], "little") * _R_INV_R % R_MOD for i in range(0, len(b), 32)]


@dataclass
class SparseR1CS:
    """cs_bn254.SparseR1CS as the reference builds it: one row per arithmetic opcode,
    qL*xa + qR*xb + qO*xc + qM*(xa*xb) + qC == 0 (sparse_r1cs.go:17)."""
    nb_public: int
    nb_secret: int
    ql: List[int]
    qr: List[int]
    qm: List[int]
    qo: List[int]
    qk: List[int]
    a: List[int]
    b: List[int]
    c: List[int]

    @property
    def nb_constraints(self) -> int:
        return len(self.ql)


def build_sparse_r1cs(acir_json: str, values: Sequence[int]):
    """ACIR JSON + dense witness values -> (SparseR1CS, public values, secret values), bug-compatible with
    backend.HandleValues (common.go:45-76) and handleArithmeticOpcode (sparse_r1cs.go:44-107)."""
    d = json.loads(acir_json)
    pubs = [int(x) for x in d["public_inputs"]]
    public_vals, secret_vals, imap = [], [], {}
    nb_public = nb_secret = 0
    for i, v in enumerate(values, start=1):
        for p in pubs:
            if i == p:
                imap[i] = nb_public
                nb_public += 1
                public_vals.append(v % R_MOD)
    for i, v in enumerate(values, start=1):
        if pubs:
            for p in pubs:
                if i != p:
                    imap[i] = nb_public + nb_secret
                    nb_secret += 1
                    secret_vals.append(v % R_MOD)
        else:
            imap[i] = nb_public + nb_secret
            nb_secret += 1
            secret_vals.append(v % R_MOD)
    cs = SparseR1CS(nb_public, nb_secret, [], [], [], [], [], [], [], [])
    for op in d["opcodes"]:
        if "Arithmetic" not in op:
            if "BlackBoxFuncCall" in op or "Directive" in op:
                continue  # no constraints (components.go:3-40, sparse_r1cs.go:36)
            raise ValueError("unknown opcode type")
        ar = op["Arithmetic"]
        mul_terms, lin = ar["mul_terms"], ar["linear_combinations"]
        xa = xb = xc = 0
        ql = qr = qo = qm = 0
        if mul_terms:
            qm = int(mul_terms[0][0], 16) % R_MOD
            xa, xb = imap.get(int(mul_terms[0][1]), 0), imap.get(int(mul_terms[0][2]), 0)
        if len(lin) == 1:
            qo, xc = int(lin[0][0], 16) % R_MOD, imap.get(int(lin[0][1]), 0)
        if len(lin) == 2:
            ql, xa = int(lin[0][0], 16) % R_MOD, imap.get(int(lin[0][1]), 0)
            qr, xb = int(lin[1][0], 16) % R_MOD, imap.get(int(lin[1][1]), 0)
        if len(lin) == 3:
            ql, xa = int(lin[0][0], 16) % R_MOD, imap.get(int(lin[0][1]), 0)
            qr, xb = int(lin[1][0], 16) % R_MOD, imap.get(int(lin[1][1]), 0)
            qo, xc = int(lin[2][0], 16) % R_MOD, imap.get(int(lin[2][1]), 0)
        for col, v in ((cs.ql, ql), (cs.qr, qr), (cs.qm, qm), (cs.qo, qo), (cs.qk, int(ar["q_c"], 16) % R_MOD),
                       (cs.a, xa), (cs.b, xb), (cs.c, xc)):
            col.append(v)
    return cs, public_vals, secret_vals


def _next_pow2_log(x: int) -> int:
    n = 0
    while (1 << n) < x:
        n += 1
    return n


def build_permutation(lro: np.ndarray) -> np.ndarray:
    """gnark buildPermutation: every position points to the previous position holding the same wire, the first
    occurrence points to the last one (vectorised: stable sort by wire id)."""
    order = np.argsort(lro, kind="stable")
    w = lro[order]
    first = np.ones(len(w), dtype=bool)
    first[1:] = w[1:] != w[:-1]
    last = np.ones(len(w), dtype=bool)
    last[:-1] = w[1:] != w[:-1]
    prev = np.empty(len(w), dtype=np.int64)
    prev[1:] = order[:-1]
    # the first occurrence of each wire points to the last occurrence of the same wire
    group_last = order[last]                       # one entry per wire group, in group order
    group_id = np.cumsum(first) - 1
    prev[first] = group_last[group_id[first]]
    perm = np.empty(len(w), dtype=np.int64)
    perm[order] = prev
    return perm


class UnsatisfiedConstraint(ValueError):
    """The solution vector violates constraint #index (what gnark's solver reports from plonk.Prove)."""

    def __init__(self, index: int):
        super().__init__("constraint #%d is not satisfied" % index)
        self.index = index


@dataclass
class Proof:
    """plonk.Proof of gnark v0.8.0 (bn254): points as 64-byte G1Affine images, scalars as 32-byte fr images."""
    blob: bytes  # 832 bytes, layout documented at b200zk_plonk_prove

    def points(self) -> List[bytes]:
        return [self.blob[64 * i: 64 * i + 64] for i in range(9)]

    def scalars(self) -> List[int]:
        return fr_from_mont(self.blob[576:])

    @staticmethod
    def _compress(pt: bytes) -> bytes:
        x = int.from_bytes(pt[:32], "little") * _R_INV_P % P_MOD
        y = int.from_bytes(pt[32:], "little") * _R_INV_P % P_MOD
        if pt == b"\0" * 64:
            return b"\x40" + b"\0" * 31
        b = bytearray(x.to_bytes(32, "big"))
        b[0] |= 0x80 if y <= (P_MOD - 1) // 2 else 0xC0
        return bytes(b)

    def to_gnark_bytes(self) -> bytes:
        """proof.WriteTo: LRO[3], Z, H[3] compressed | BatchedProof.H | u32 7 | 7 fr | ZShifted.H | fr  (548 B)."""
        p = self.points()
        s = self.scalars()
        out = b"".join(self._compress(x) for x in p[:7]) + self._compress(p[7])
        out += (7).to_bytes(4, "big") + b"".join(v.to_bytes(32, "big") for v in s[:7])
        out += self._compress(p[8]) + s[7].to_bytes(32, "big")
        return out


class ProvingKey:
    """plonk.ProvingKey resident on the device (selectors, permutation polynomials, their Lagrange-coset forms, SRS)."""

    def __init__(self):
        self.handle = None

    @classmethod
    def Setup(cls, cs: SparseR1CS, srs: SRS, ctx: Optional[Context] = None) -> "ProvingKey":
        """plonk.Setup(spr, srs) through b200zk_plonk_setup_r1cs: the row layout [placeholders | constraints | padding]
        and gnark's permutation are built inside the library (C++); this wrapper only marshals the columns."""
        self = cls()
        self.ctx = ctx or srs.ctx
        self.srs = srs
        size_system = cs.nb_constraints + cs.nb_public
        self.log2n = max(_next_pow2_log(size_system), 1)
        self.n = 1 << self.log2n
        self.log2n_big = max(_next_pow2_log((8 if size_system < 6 else 4) * size_system), self.log2n + 2)
        self.nb_public = cs.nb_public
        self.nb_wires = max(cs.nb_public + cs.nb_secret, 1)
        cols = [np.ascontiguousarray(fr_to_mont(v)) if len(v) else np.zeros(32, dtype=np.uint8)
                for v in (cs.ql, cs.qr, cs.qm, cs.qo, cs.qk)]
        wires = [np.ascontiguousarray(np.asarray(v, dtype=np.uint32)) if len(v) else np.zeros(1, dtype=np.uint32)
                 for v in (cs.a, cs.b, cs.c)]
        lib, h = self.ctx.lib, self.ctx.handle
        out = C.c_void_p()
        rc = lib.b200zk_plonk_setup_r1cs(h, srs.handle, cs.nb_public, cs.nb_secret, cs.nb_constraints,
                                         *[c.ctypes.data for c in cols], *[w.ctypes.data for w in wires], C.byref(out))
        _lib.check(h, rc)
        self.handle = out
        vk = np.zeros(8 * 64, dtype=np.uint8)
        _lib.check(h, lib.b200zk_plonk_vk(h, self.handle, vk.ctypes.data))
        self.vk_points = [vk[64 * i: 64 * i + 64].tobytes() for i in range(8)]
        return self

    @classmethod
    def SetupPy(cls, cs: SparseR1CS, srs: SRS, ctx: Optional[Context] = None) -> "ProvingKey":
        """Same as Setup with the row layout / permutation built in numpy (cross-check of the C++ builder)."""
        self = cls()
        self.ctx = ctx or srs.ctx
        self.srs = srs
        size_system = cs.nb_constraints + cs.nb_public
        self.log2n = max(_next_pow2_log(size_system), 1)
        n = 1 << self.log2n
        self.log2n_big = _next_pow2_log((8 if size_system < 6 else 4) * size_system)
        self.log2n_big = max(self.log2n_big, self.log2n + 2)
        self.n = n
        self.nb_public = cs.nb_public
        self.nb_wires = max(cs.nb_public + cs.nb_secret, 1)
        off = cs.nb_public

        def column(vals, placeholder):
            col = [placeholder] * off + list(vals) + [0] * (n - off - len(vals))
            return np.ascontiguousarray(fr_to_mont(col))

        ql = column(cs.ql, R_MOD - 1)
        qr, qm, qo, qk = (column(v, 0) for v in (cs.qr, cs.qm, cs.qo, cs.qk))
        lro = np.zeros(3 * n, dtype=np.uint32)
        lro[:off] = np.arange(off, dtype=np.uint32)
        k = cs.nb_constraints
        lro[off:off + k] = np.asarray(cs.a, dtype=np.uint32)
        lro[n + off:n + off + k] = np.asarray(cs.b, dtype=np.uint32)
        lro[2 * n + off:2 * n + off + k] = np.asarray(cs.c, dtype=np.uint32)
        perm = build_permutation(lro)
        self.lro, self.permutation = lro, perm
        return self._finish(ql, qr, qm, qo, qk, perm, lro)

    @classmethod
    def SetupRaw(cls, srs: SRS, log2n: int, log2n_big: int, nb_public: int, nb_wires: int, ql, qr, qm, qo, qk,
                 lro: np.ndarray, ctx: Optional[Context] = None) -> "ProvingKey":
        """Setup from ready-made Lagrange columns (n*32-byte Montgomery images) — for large synthetic circuits."""
        self = cls()
        self.ctx = ctx or srs.ctx
        self.srs = srs
        self.log2n, self.log2n_big, self.n = log2n, log2n_big, 1 << log2n
        self.nb_public, self.nb_wires = nb_public, nb_wires
        perm = build_permutation(lro)
        self.lro, self.permutation = lro, perm
        return self._finish(*(np.ascontiguousarray(x) for x in (ql, qr, qm, qo, qk)), perm, lro)

    def _finish(self, ql, qr, qm, qo, qk, perm, lro) -> "ProvingKey":
        lib, h = self.ctx.lib, self.ctx.handle
        out = C.c_void_p()
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        lro = np.ascontiguousarray(lro, dtype=np.uint32)
        rc = lib.b200zk_plonk_setup(h, self.srs.handle, self.log2n, self.log2n_big, self.nb_public, self.nb_wires,
                                    ql.ctypes.data, qr.ctypes.data, qm.ctypes.data, qo.ctypes.data, qk.ctypes.data,
                                    perm.ctypes.data, lro.ctypes.data, C.byref(out))
        _lib.check(h, rc)
        self.handle = out
        vk = np.zeros(8 * 64, dtype=np.uint8)
        _lib.check(h, lib.b200zk_plonk_vk(h, self.handle, vk.ctypes.data))
        self.vk_points = [vk[64 * i: 64 * i + 64].tobytes() for i in range(8)]  # S0,S1,S2,Ql,Qr,Qm,Qo,Qk
        return self

    def poly(self, which: int) -> bytes:
        """Ql,Qr,Qm,Qo,CQk (canonical), LQk (Lagrange), S1,S2,S3 (canonical) — for ProvingKey.WriteTo."""
        out = np.zeros(self.n * 32, dtype=np.uint8)
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_plonk_pk_poly(self.ctx.handle, self.handle, which, out.ctypes.data))
        return out.tobytes()

    # ---- multi-GPU: one key per GPU, the ranks prove together (b200zk_plonk_join) -------------------------
    def arena(self):
        """(device pointer, bytes) of the key's arena — what the other ranks map"""
        base, nbytes = C.c_void_p(), C.c_size_t()
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_plonk_arena(self.ctx.handle, self.handle, C.byref(base), C.byref(nbytes)))
        return base.value, nbytes.value

    def Join(self, group=None) -> None:
        """Join the ranks of a torch.distributed group (one process per GPU of one box) into one prover: every rank
        holds a key set up for the same circuit; the arenas are mapped into each other through CUDA IPC handles
        exchanged once (all_gather), after which no collective is on the data path.  Afterwards every rank calls
        Prove (ranks > 0 may pass None for solution and blinding) and gets the same proof."""
        import torch
        import torch.distributed as dist

        lib, h = self.ctx.lib, self.ctx.handle
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        base, _ = self.arena()
        handle = (C.c_ubyte * 64)()
        _lib.check(h, lib.b200zk_ipc_export(h, C.c_void_p(base), handle))
        dev = "cuda:%d" % self.ctx.device
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        ptrs = (C.c_void_p * world)()
        self._imported = []
        for k in range(world):
            if k == rank:
                ptrs[k] = base
            else:
                raw = (C.c_ubyte * 64)(*allh[k].cpu().tolist())
                out = C.c_void_p()
                _lib.check(h, lib.b200zk_ipc_import(h, raw, C.byref(out)))
                ptrs[k] = out.value
                self._imported.append(out)
        _lib.check(h, lib.b200zk_plonk_join(h, self.handle, rank, world, ptrs))
        self._group = group
        self.rank, self.world = rank, world

    @staticmethod
    def JoinLocal(keys: Sequence["ProvingKey"]) -> None:
        """The same for keys that live in ONE process (one context per key, on one or several GPUs): the arena pointers
        are valid as they are.  Every key must then be proved from its own host thread (the ranks wait for each other
        on the device).  Keys that share one DEVICE (tests) need CUDA_MODULE_LOADING=EAGER and one hardware queue per
        stream (CUDA_DEVICE_MAX_CONNECTIONS=32): a waiting rank must never be behind a context-wide synchronisation."""
        world = len(keys)
        ptrs = (C.c_void_p * world)(*[k.arena()[0] for k in keys])
        for r, k in enumerate(keys):
            _lib.check(k.ctx.handle, k.ctx.lib.b200zk_plonk_join(k.ctx.handle, k.handle, r, world, ptrs))
            k.rank, k.world, k._imported, k._group = r, world, [], None

    def Leave(self) -> None:
        """back to single-GPU proving (collective over the group when the key was joined with Join)"""
        if getattr(self, "world", 1) <= 1:
            return
        lib, h = self.ctx.lib, self.ctx.handle
        _lib.check(h, lib.b200zk_plonk_leave(h, self.handle))
        if self._imported:
            import torch.distributed as dist

            dist.barrier(group=self._group)   # nobody unmaps an arena a peer may still be reading
            for p in self._imported:
                lib.b200zk_ipc_close(h, p)
            dist.barrier(group=self._group)
        self._imported = []
        self.world, self.rank = 1, 0

    def Prove(self, solution, blinding) -> Proof:
        """plonk.Prove: solution = every wire's value (public wires first), blinding = 9 fr.SetRandom draws
        (Montgomery images, order L,L,R,R,O,O,Z,Z,Z).  Joined keys: ranks > 0 may pass None for both."""
        out = np.zeros(832, dtype=np.uint8)
        if solution is None:
            assert getattr(self, "world", 1) > 1 and self.rank != 0, "only ranks > 0 of a joined key take the solution from rank 0"
            rc = self.ctx.lib.b200zk_plonk_prove(self.ctx.handle, self.handle, None, None, out.ctypes.data)
            if rc == _lib.ERR_UNSATISFIED:
                raise UnsatisfiedConstraint(self.ctx.lib.b200zk_plonk_unsatisfied_row(self.handle) - self.nb_public)
            _lib.check(self.ctx.handle, rc)
            return Proof(out.tobytes())
        sol = np.ascontiguousarray(np.frombuffer(bytes(solution), dtype=np.uint8) if not isinstance(solution, np.ndarray) else solution)
        bl = np.ascontiguousarray(np.frombuffer(bytes(blinding), dtype=np.uint8) if not isinstance(blinding, np.ndarray) else blinding)
        assert sol.nbytes == self.nb_wires * 32 and bl.nbytes == 9 * 32
        rc = self.ctx.lib.b200zk_plonk_prove(self.ctx.handle, self.handle, sol.ctypes.data, bl.ctypes.data, out.ctypes.data)
        if rc == _lib.ERR_UNSATISFIED:  # plonk.Prove returns spr.Solve's error before committing to anything
            row = self.ctx.lib.b200zk_plonk_unsatisfied_row(self.handle)
            raise UnsatisfiedConstraint(row - self.nb_public)
        _lib.check(self.ctx.handle, rc)
        return Proof(out.tobytes())

    def SetSolutionMap(self, src, nb_values: int) -> None:
        """BuildWitnesses as a gather: wire k <- values[src[k]] (once per key, for ProveHex)."""
        src = np.ascontiguousarray(np.asarray(src, dtype=np.uint32))
        assert src.size == self.nb_wires
        rc = self.ctx.lib.b200zk_plonk_set_solution_map(self.ctx.handle, self.handle, src.ctypes.data, int(nb_values))
        _lib.check(self.ctx.handle, rc)

    def ProveHex(self, values_hex: bytes, nb_values: int, blinding) -> Proof:
        """plonk.Prove fed with the hex text PlonkProveWithPK receives (64 characters per value, no count prefix):
        decoding, Montgomery conversion and the witness gather run on the device."""
        assert len(values_hex) == 64 * nb_values
        bl = np.ascontiguousarray(np.frombuffer(bytes(blinding), dtype=np.uint8) if not isinstance(blinding, np.ndarray) else blinding)
        out = np.zeros(832, dtype=np.uint8)
        rc = self.ctx.lib.b200zk_plonk_prove_hex(self.ctx.handle, self.handle, values_hex, nb_values, bl.ctypes.data, out.ctypes.data)
        if rc == _lib.ERR_UNSATISFIED:
            raise UnsatisfiedConstraint(self.ctx.lib.b200zk_plonk_unsatisfied_row(self.handle) - self.nb_public)
        _lib.check(self.ctx.handle, rc)
        return Proof(out.tobytes())

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            if getattr(self, "world", 1) > 1 and not getattr(self, "_imported", None):
                self.ctx.lib.b200zk_plonk_leave(self.ctx.handle, self.handle)
            self.ctx.lib.b200zk_plonk_pk_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
