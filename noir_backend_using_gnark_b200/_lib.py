"""ctypes binding of libb200zk.so (include/b200zk.h).  Fails loudly when the library is missing: there is no
CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libb200zk.so"
HEADER = PKG.parent / "include" / "b200zk.h"

DIF, DIT = 0, 1
ERR_UNSATISFIED = -6
MAX_LOG2N = 28


class B200zkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200zk error %d: %s" % (code, msg))
        self.code = code


_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/b200zk.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        # not a fallback: the same CUDA library, compiled on the spot when the prebuilt .so did not travel
        try:
            from . import build as _build

            _build.build()
        except Exception as e:
            raise ImportError(
                "%s is missing and could not be built (%s): run `python -m noir_backend_using_gnark_b200.build` "
                "(this package has no CPU fallback)" % (LIB_PATH, e)
            )
    lib = C.CDLL(str(LIB_PATH))
    vp, sz, u, i = C.c_void_p, C.c_size_t, C.c_uint, C.c_int
    sig = {
        "b200zk_device_count": (i, []),
        "b200zk_init": (i, [i, C.POINTER(vp)]),
        "b200zk_destroy": (None, [vp]),
        "b200zk_strerror": (C.c_char_p, [i]),
        "b200zk_last_cuda_error": (C.c_char_p, [vp]),
        "b200zk_stream": (vp, [vp]),
        "b200zk_sync": (i, [vp]),
        "b200zk_launch_count": (C.c_uint64, [vp]),
        "b200zk_ntt": (i, [vp, vp, u, i, i, i]),
        "b200zk_ntt_dev": (i, [vp, vp, u, i, i, i]),
        "b200zk_ntt_dist_half_dev": (i, [vp, vp, vp, u, u, u, u, i, i, i, i]),
        "b200zk_ntt_set_radix2": (i, [vp, i]),
        "b200zk_ntt_dist_half0_p2p_dev": (i, [vp, vp, C.POINTER(vp), u, u, u, u, i, i, i]),
        "b200zk_dev_alloc": (i, [vp, sz, C.POINTER(vp)]),
        "b200zk_dev_free": (i, [vp, vp]),
        "b200zk_host_alloc": (i, [vp, sz, C.POINTER(vp)]),
        "b200zk_host_free": (i, [vp, vp]),
        "b200zk_ipc_export": (i, [vp, vp, vp]),
        "b200zk_ipc_import": (i, [vp, vp, C.POINTER(vp)]),
        "b200zk_ipc_close": (i, [vp, vp]),
        "b200zk_bit_reverse": (i, [vp, vp, u]),
        "b200zk_bit_reverse_dev": (i, [vp, vp, u]),
        "b200zk_bases_upload": (i, [vp, vp, sz, C.POINTER(vp)]),
        "b200zk_bases_wrap_dev": (i, [vp, vp, sz, C.POINTER(vp)]),
        "b200zk_srs_generate": (i, [vp, vp, sz, sz, C.POINTER(vp)]),
        "b200zk_bases_download": (i, [vp, vp, sz, sz, vp]),
        "b200zk_bases_upload_compressed": (i, [vp, vp, sz, C.POINTER(vp)]),
        "b200zk_bases_download_compressed": (i, [vp, vp, sz, sz, vp]),
        "b200zk_bases_precompute": (i, [vp, vp, i]),
        "b200zk_bases_free": (None, [vp, vp]),
        "b200zk_bases_len": (sz, [vp]),
        "b200zk_msm_g1": (i, [vp, vp, vp, sz, vp]),
        "b200zk_msm_g1_shard": (i, [vp, vp, sz, vp, sz, vp]),
        "b200zk_msm_windows": (i, [vp, vp, sz]),
        "b200zk_msm_set_host_chunks": (i, [vp, i]),
        "b200zk_msm_set_small_path": (i, [vp, i]),
        "b200zk_msm_set_pair_rounds": (i, [vp, i]),
        "b200zk_msm_set_scatter_passes": (i, [vp, i]),
        "b200zk_msm_set_reduce_chunk": (i, [vp, i]),
        "b200zk_msm_g1_dev": (i, [vp, vp, sz, vp, sz, vp, i]),
        "b200zk_g1_sum_dev": (i, [vp, vp, sz, vp]),
        "b200zk_msm_set_window": (i, [vp, i]),
        "b200zk_plonk_setup": (i, [vp, vp, u, u, u, u, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "b200zk_plonk_setup_r1cs": (i, [vp, vp, u, u, sz, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]),
        "b200zk_plonk_pk_free": (None, [vp, vp]),
        "b200zk_plonk_vk": (i, [vp, vp, vp]),
        "b200zk_plonk_set_commit_hook": (i, [vp, vp, vp]),
        "b200zk_plonk_pk_poly": (i, [vp, vp, i, vp]),
        "b200zk_plonk_prove": (i, [vp, vp, vp, vp, vp]),
        "b200zk_plonk_unsatisfied_row": (C.c_longlong, [vp]),
        "b200zk_plonk_arena": (i, [vp, vp, C.POINTER(vp), C.POINTER(sz)]),
        "b200zk_plonk_join": (i, [vp, vp, u, u, C.POINTER(vp)]),
        "b200zk_plonk_leave": (i, [vp, vp]),
        "b200zk_plonk_set_commit_lanes": (i, [vp, i]),
        "b200zk_plonk_set_solution_map": (i, [vp, vp, vp, sz]),
        "b200zk_plonk_prove_hex": (i, [vp, vp, vp, sz, vp, vp]),
        "b200zk_microbench": (i, [vp, i, C.POINTER(C.c_double)]),
        "b200zk_profile_enable": (i, [vp, i]),
        "b200zk_profile_read": (i, [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._signatures = sig
    _lib = lib
    return lib


def check(ctx, rc: int) -> None:
    if rc != 0:
        lib = load()
        msg = lib.b200zk_strerror(rc).decode()
        if ctx:
            detail = lib.b200zk_last_cuda_error(ctx).decode()
            if detail:
                msg += " (" + detail + ")"
        raise B200zkError(rc, msg)
