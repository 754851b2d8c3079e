"""ctypes loader for the C oracle (oracle/bn254_ref.c).  TEST INFRASTRUCTURE ONLY — see the header of that file."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libbn254_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "bn254_ref.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        # -march=native is avoided: the .so is built here and travels to a different host
        cmd = ["gcc", "-O3", "-fPIC", "-pthread", "-shared", "-o", str(LIB), str(src)]
        subprocess.run(cmd, check=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        vp, sz, u, i = C.c_void_p, C.c_size_t, C.c_uint, C.c_int
        lib.oracle_ntt.restype = i
        lib.oracle_ntt.argtypes = [vp, u, i, i, i, i]
        lib.oracle_bit_reverse.argtypes = [vp, u]
        lib.oracle_domain_release.argtypes = [u]
        lib.oracle_msm.restype = i
        lib.oracle_msm.argtypes = [vp, vp, sz, vp, i, i]
        lib.oracle_g1_arith_progression.restype = i
        lib.oracle_g1_arith_progression.argtypes = [vp, vp, sz, vp]
        lib.oracle_g1_add_affine.argtypes = [vp, vp, vp]
        lib.oracle_random_fr.argtypes = [vp, sz, C.c_uint64]
        lib.oracle_g1_mul_gen_batch.argtypes = [vp, sz, vp, i]
        lib.oracle_plonk_prove.restype = i
        lib.oracle_plonk_prove.argtypes = [u, u, u, u, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp, vp, i, vp]
        for name in ("oracle_fr_mul", "oracle_fp_mul", "oracle_fr_add", "oracle_fr_sub"):
            getattr(lib, name).argtypes = [vp, vp, vp]
        for name in ("oracle_fr_inv", "oracle_fp_inv"):
            getattr(lib, name).argtypes = [vp, vp]
        _lib = lib
    return _lib


def ncores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _buf(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return b
    return np.frombuffer(bytearray(b), dtype=np.uint8)


def ntt(data, log2n: int, inverse: bool, decimation: int, coset: bool, nthreads: int = 1) -> bytes:
    a = _buf(data).copy()
    assert a.nbytes == (1 << log2n) * 32
    rc = load().oracle_ntt(a.ctypes.data, log2n, int(inverse), decimation, int(coset), nthreads)
    assert rc == 0
    return a.tobytes()


def ntt_inplace(a: np.ndarray, log2n: int, inverse: bool, decimation: int, coset: bool, nthreads: int = 1) -> None:
    rc = load().oracle_ntt(a.ctypes.data, log2n, int(inverse), decimation, int(coset), nthreads)
    assert rc == 0


def bit_reverse(data, log2n: int) -> bytes:
    a = _buf(data).copy()
    load().oracle_bit_reverse(a.ctypes.data, log2n)
    return a.tobytes()


def msm(points, scalars, n: int | None = None, nthreads: int = 1, c: int = 0) -> bytes:
    p = _buf(points)
    s = _buf(scalars)
    if n is None:
        n = s.nbytes // 32
    assert p.nbytes >= n * 64 and s.nbytes >= n * 32
    out = np.zeros(64, dtype=np.uint8)
    rc = load().oracle_msm(p.ctypes.data, s.ctypes.data, n, out.ctypes.data, nthreads, c)
    assert rc == 0
    return out.tobytes()


def random_fr(n: int, seed: int) -> np.ndarray:
    """n uniform fr elements as the in-memory (Montgomery) byte image, uint8 array of n*32."""
    out = np.zeros(n * 32, dtype=np.uint8)
    load().oracle_random_fr(out.ctypes.data, n, seed & 0xFFFFFFFFFFFFFFFF)
    return out


def g1_arith_progression(first_affine: bytes, step_affine: bytes, n: int) -> np.ndarray:
    """P_i = first + i*step (affine, 64 B each)."""
    f = _buf(first_affine)
    s = _buf(step_affine)
    out = np.zeros(n * 64, dtype=np.uint8)
    rc = load().oracle_g1_arith_progression(f.ctypes.data, s.ctypes.data, n, out.ctypes.data)
    assert rc == 0
    return out


def g1_mul_gen_batch(scalars_mont, n: int | None = None, nthreads: int | None = None) -> np.ndarray:
    """[s_i * G] for Montgomery-form scalars; (n*64,) uint8."""
    s = _buf(scalars_mont)
    if n is None:
        n = s.nbytes // 32
    out = np.zeros(n * 64, dtype=np.uint8)
    load().oracle_g1_mul_gen_batch(s.ctypes.data, n, out.ctypes.data, nthreads or ncores())
    return out
