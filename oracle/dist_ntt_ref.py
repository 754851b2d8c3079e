"""Big-int checker for ONE local half of the multi-GPU four-step NTT (TEST INFRASTRUCTURE ONLY).

Restates, on logical indices, what b200zk_ntt_dist_half_dev must do to one rank's shard; used by the gloo CPU test
(tests/test_dist_cpu.py) to validate the decomposition + all-to-all layout arithmetic of
noir_backend_using_gnark_b200/dist_ntt.py without a GPU, and by the 2-GPU test as the per-half oracle."""
from __future__ import annotations

import numpy as np

from . import bn254 as o


def _exchange_indices(lay, rank: int) -> np.ndarray:
    """logical index of Z[peer][r_lo][c_lo] on row-block rank `rank`"""
    peer = np.arange(lay.world, dtype=np.int64)[:, None, None]
    r_lo = np.arange(lay.R_loc, dtype=np.int64)[None, :, None]
    c_lo = np.arange(lay.C_loc, dtype=np.int64)[None, None, :]
    return ((rank * lay.R_loc + r_lo) * lay.C + peer * lay.C_loc + c_lo).reshape(-1)


def half(lay, rank: int, src_bytes: bytes, half_idx: int, inverse: int, decimation: int, coset: int) -> bytes:
    n = lay.log2n
    dom = o.Domain(1 << n)
    w = dom.generator_inv if inverse else dom.generator
    dit = decimation == o.DIT
    high_half = (half_idx == 1) if dit else (half_idx == 0)
    col = lay.column_block_indices(rank)
    row = lay.row_block_indices(rank)
    exch = _exchange_indices(lay, rank)
    if high_half:
        src_idx = dst_idx = col
        stages = list(range(lay.log2c, n))
    else:
        stages = list(range(0, lay.log2c))
        src_idx, dst_idx = (row, exch) if dit else (exch, row)
    vals = dict(zip(src_idx.tolist(), o.fr_from_mont_bytes(src_bytes)))
    brev = lambda i: o.bit_reverse_index(i, n)
    if not inverse and coset and half_idx == 0:
        for i in vals:
            e = brev(i) if dit else i
            vals[i] = vals[i] * pow(o.FR_COSET_GEN, e, o.R_MOD) % o.R_MOD
    order = stages if dit else stages[::-1]
    for hbit in order:
        h = 1 << hbit
        for i in list(vals):
            if i & h:
                continue
            j = i & (h - 1)
            tw = pow(w, j << (n - 1 - hbit), o.R_MOD)
            u, v = vals[i], vals[i + h]
            if dit:
                v = v * tw % o.R_MOD
                vals[i], vals[i + h] = (u + v) % o.R_MOD, (u - v) % o.R_MOD
            else:
                vals[i], vals[i + h] = (u + v) % o.R_MOD, (u - v) * tw % o.R_MOD
    if inverse and half_idx == 1:
        ninv = dom.cardinality_inv
        gi = dom.fr_multiplicative_gen_inv
        for i in vals:
            f = ninv
            if coset:
                e = i if dit else brev(i)
                f = f * pow(gi, e, o.R_MOD) % o.R_MOD
            vals[i] = vals[i] * f % o.R_MOD
    return o.fr_to_mont_bytes(vals[i] for i in dst_idx.tolist())
