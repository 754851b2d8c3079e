"""Host-side mirror of the reference's arithmetic interface for the PLONK hot path, over the C ABI.

Names follow gnark-crypto v0.9.1 (the module /root/reference/gnark_backend_ffi/go.mod:5 pins and
/root/reference/gnark_backend_ffi/backend/plonk/plonk.go:21,67 reaches through gnark):
  fft.NewDomain / Domain.FFT / Domain.FFTInverse / fft.BitReverse   ->  Domain(...).FFT / .FFTInverse, BitReverse
  kzg.SRS.G1 + (*G1Affine).MultiExp / kzg.Commit                     ->  SRS(...), MultiExp, Commit
Byte layouts are gnark's in-memory ones (see include/b200zk.h), carried as `bytes` / numpy uint8 on the host and
as torch uint8 CUDA tensors on the device.  Every call goes through libb200zk.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import DIF, DIT, B200zkError  # noqa: F401


def _host_ptr(buf) -> tuple[int, object]:
    """(address, keep-alive) for a writable/readable contiguous host buffer."""
    if isinstance(buf, np.ndarray):
        assert buf.flags["C_CONTIGUOUS"]
        return buf.ctypes.data, buf
    if isinstance(buf, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(buf, dtype=np.uint8)
        return arr.ctypes.data, arr
    if hasattr(buf, "data_ptr"):  # torch CPU tensor (pinned or not)
        assert not buf.is_cuda and buf.is_contiguous()
        return buf.data_ptr(), buf
    raise TypeError("unsupported host buffer %r" % type(buf))


class Context:
    """One context per GPU (b200zk_init).  `device` is the CUDA ordinal."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.b200zk_init(int(device), C.byref(h))
        _lib.check(None, rc)
        self.handle = h
        self.device = int(device)

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.b200zk_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def sync(self) -> None:
        _lib.check(self.handle, self.lib.b200zk_sync(self.handle))

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200zk_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200zk_launch_count(self.handle))

    def microbench(self, which: int) -> float:
        """ops/s of the integer-pipe microbenchmarks (0: IMAD.WIDE, 1: fp mul, 2: fr mul)."""
        v = C.c_double()
        _lib.check(self.handle, self.lib.b200zk_microbench(self.handle, which, C.byref(v)))
        return v.value

    PHASES = ("msm_digits", "msm_scan", "msm_scatter", "msm_accumulate", "msm_big", "msm_reduce", "msm_final",
              "ntt_pass")

    def profile(self, on: bool) -> None:
        _lib.check(self.handle, self.lib.b200zk_profile_enable(self.handle, int(on)))

    def profile_read(self) -> dict:
        """{phase: (total_ms, count)} since the last read (synchronises the context stream)."""
        ms = (C.c_double * 8)()
        cnt = (C.c_uint64 * 8)()
        _lib.check(self.handle, self.lib.b200zk_profile_read(self.handle, ms, cnt, 8))
        return {name: (ms[k], int(cnt[k])) for k, name in enumerate(self.PHASES)}

    def torch_stream(self):
        import torch

        return torch.cuda.ExternalStream(self.stream_ptr, device=self.device)

    def after_torch(self) -> None:
        """Device-tensor entry points: the library works on its own non-blocking stream, so the work already queued on
        torch's current stream (the producers of the tensors passed in) is ordered before it here.  Results are ready
        after `sync()`, or for torch after `before_torch()`."""
        import torch

        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream_ptr:
            self.torch_stream().wait_stream(cur)

    def before_torch(self) -> None:
        """torch's current stream waits for everything enqueued on the library's stream so far"""
        import torch

        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != self.stream_ptr:
            cur.wait_stream(self.torch_stream())


_default: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]


# ------------------------------------------------------------------------------------------------------
# fft
# ------------------------------------------------------------------------------------------------------
def _log2_ceil(m: int) -> int:
    n = 0
    while (1 << n) < m:
        n += 1
    return n


class Domain:
    """fft.NewDomain(m): Cardinality = next power of two >= m."""

    def __init__(self, m: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.log2n = _log2_ceil(int(m))
        if self.log2n > _lib.MAX_LOG2N:
            raise B200zkError(-3, "domain larger than 2^28")
        self.Cardinality = 1 << self.log2n

    def _run(self, a, inverse: int, decimation: int, coset: bool):
        lib, h = self.ctx.lib, self.ctx.handle
        if hasattr(a, "is_cuda") and a.is_cuda:
            assert a.is_contiguous() and a.numel() * a.element_size() == self.Cardinality * 32
            assert a.device.index == self.ctx.device
            self.ctx.after_torch()
            rc = lib.b200zk_ntt_dev(h, a.data_ptr(), self.log2n, inverse, decimation, int(bool(coset)))
            _lib.check(h, rc)
            return a
        if isinstance(a, bytes):
            a = bytearray(a)
        ptr, keep = _host_ptr(a)
        assert keep.nbytes if hasattr(keep, "nbytes") else True
        rc = lib.b200zk_ntt(h, ptr, self.log2n, inverse, decimation, int(bool(coset)))
        _lib.check(h, rc)
        return a

    def FFT(self, a, decimation: int, coset: bool = False):
        """domain.FFT(a, decimation, coset) in place; returns `a` (a new bytearray if `a` was immutable bytes)."""
        return self._run(a, 0, decimation, coset)

    def FFTInverse(self, a, decimation: int, coset: bool = False):
        return self._run(a, 1, decimation, coset)


def BitReverse(a, ctx: Optional[Context] = None):
    """fft.BitReverse(a) in place."""
    ctx = ctx or default_context()
    lib, h = ctx.lib, ctx.handle
    if hasattr(a, "is_cuda") and a.is_cuda:
        n = a.numel() * a.element_size() // 32
        log2n = _log2_ceil(n)
        assert 1 << log2n == n
        ctx.after_torch()
        _lib.check(h, lib.b200zk_bit_reverse_dev(h, a.data_ptr(), log2n))
        return a
    if isinstance(a, bytes):
        a = bytearray(a)
    ptr, keep = _host_ptr(a)
    n = len(a) // 32 if not isinstance(a, np.ndarray) else a.nbytes // 32
    log2n = _log2_ceil(n)
    assert 1 << log2n == n
    _lib.check(h, lib.b200zk_bit_reverse(h, ptr, log2n))
    return a


# ------------------------------------------------------------------------------------------------------
# kzg / multiexp
# ------------------------------------------------------------------------------------------------------
class SRS:
    """The G1 half of kzg.SRS: the MSM bases, resident on the device (uploaded once)."""

    def __init__(self, g1_affine, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        lib, h = self.ctx.lib, self.ctx.handle
        out = C.c_void_p()
        if hasattr(g1_affine, "is_cuda") and g1_affine.is_cuda:
            n = g1_affine.numel() * g1_affine.element_size() // 64
            self._keep = g1_affine
            rc = lib.b200zk_bases_wrap_dev(h, g1_affine.data_ptr(), n, C.byref(out))
        else:
            ptr, keep = _host_ptr(g1_affine)
            n = (keep.nbytes if hasattr(keep, "nbytes") else len(g1_affine)) // 64
            rc = lib.b200zk_bases_upload(h, ptr, n, C.byref(out))
        _lib.check(h, rc)
        self.handle = out
        self.n = n

    @classmethod
    def NewSRS(cls, size: int, alpha_mont: bytes, ctx: Optional[Context] = None, first: int = 0) -> "SRS":
        """kzg.NewSRS(first+size, alpha).G1[first:] = [alpha^(first+i) * G], generated on the device;
        alpha as a Montgomery fr.Element."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        out = C.c_void_p()
        ptr, keep = _host_ptr(alpha_mont)
        rc = self.ctx.lib.b200zk_srs_generate(self.ctx.handle, ptr, int(first), int(size), C.byref(out))
        _lib.check(self.ctx.handle, rc)
        self.handle = out
        self.n = int(size)
        return self

    @classmethod
    def FromCompressed(cls, compressed: bytes, ctx: Optional[Context] = None) -> "SRS":
        """kzg.SRS.ReadFrom's G1 part: n x 32-byte compressed points (G1Affine.Bytes()), decompressed on the device."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        ptr, keep = _host_ptr(compressed)
        n = (keep.nbytes if hasattr(keep, "nbytes") else len(compressed)) // 32
        out = C.c_void_p()
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_bases_upload_compressed(self.ctx.handle, ptr, n, C.byref(out)))
        self.handle = out
        self.n = n
        return self

    def download_compressed(self, first: int = 0, n: Optional[int] = None) -> bytes:
        """kzg.SRS.WriteTo's G1 part (without the u32 length prefix): compressed on the device."""
        n = self.n - first if n is None else n
        out = np.zeros(max(n, 1) * 32, dtype=np.uint8)
        rc = self.ctx.lib.b200zk_bases_download_compressed(self.ctx.handle, self.handle, first, n, out.ctypes.data)
        _lib.check(self.ctx.handle, rc)
        return out[: n * 32].tobytes()

    def precompute(self, c: int = 0) -> "SRS":
        """One-time window-multiple table for these (static) bases: faster MSMs, W times the memory."""
        _lib.check(self.ctx.handle, self.ctx.lib.b200zk_bases_precompute(self.ctx.handle, self.handle, int(c)))
        return self

    def windows(self, n: Optional[int] = None) -> int:
        """Bucket additions per point an MSM of n of these bases performs (measurement aid)."""
        w = self.ctx.lib.b200zk_msm_windows(self.ctx.handle, self.handle, self.n if n is None else n)
        if w < 0:
            _lib.check(self.ctx.handle, w)
        return w

    def download(self, first: int = 0, n: Optional[int] = None) -> bytes:
        n = self.n - first if n is None else n
        out = np.zeros(max(n, 1) * 64, dtype=np.uint8)
        rc = self.ctx.lib.b200zk_bases_download(self.ctx.handle, self.handle, first, n, out.ctypes.data)
        _lib.check(self.ctx.handle, rc)
        return out[: n * 64].tobytes()

    def close(self) -> None:
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.b200zk_bases_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return self.n


def MultiExp(srs: SRS, scalars, n: Optional[int] = None, first_base: int = 0, out=None, partial: bool = False):
    """(*G1Affine).MultiExp(srs.G1[first_base:first_base+n], scalars[:n]).

    Host scalars (bytes / numpy): returns the 64-byte canonical affine result as bytes.
    Device scalars (torch CUDA uint8): asynchronous; writes into / returns a CUDA tensor of 64 bytes
    (128 bytes X,Y,ZZ,ZZZ if partial=True, for cross-GPU combination with SumPartials)."""
    ctx = srs.ctx
    lib, h = ctx.lib, ctx.handle
    if hasattr(scalars, "is_cuda") and scalars.is_cuda:
        import torch

        if n is None:
            n = scalars.numel() * scalars.element_size() // 32
        if out is None:
            out = torch.empty(128 if partial else 64, dtype=torch.uint8, device=scalars.device)
        ctx.after_torch()
        rc = lib.b200zk_msm_g1_dev(h, srs.handle, first_base, scalars.data_ptr(), n, out.data_ptr(), int(partial))
        _lib.check(h, rc)
        return out
    assert first_base == 0 and not partial
    ptr, keep = _host_ptr(scalars)
    if n is None:
        n = (keep.nbytes if hasattr(keep, "nbytes") else len(scalars)) // 32
    res = np.zeros(64, dtype=np.uint8)
    rc = lib.b200zk_msm_g1(h, srs.handle, ptr, n, res.ctypes.data)
    _lib.check(h, rc)
    return res.tobytes()


def MultiExpShard(srs: SRS, scalars_host, n: Optional[int] = None, first_base: int = 0, out=None):
    """One rank's share of a point-range-sharded MultiExp, fed from HOST scalars (b200zk_msm_g1_shard): the chunked
    PCIe copy runs under the bucket work; the 128-byte extended-Jacobian partial stays on the device (CUDA tensor) for
    the cross-GPU all-gather + SumPartials."""
    import torch

    ctx = srs.ctx
    ptr, keep = _host_ptr(scalars_host)
    if n is None:
        n = (keep.nbytes if hasattr(keep, "nbytes") else len(scalars_host)) // 32
    if out is None:
        out = torch.empty(128, dtype=torch.uint8, device="cuda:%d" % ctx.device)
    _lib.check(ctx.handle, ctx.lib.b200zk_msm_g1_shard(ctx.handle, srs.handle, first_base, ptr, n, out.data_ptr()))
    return out


def Commit(p, srs: SRS):
    """kzg.Commit(p, srs): MultiExp of the polynomial's coefficients against srs.G1[:len(p)]."""
    return MultiExp(srs, p)


def SumPartials(ctx: Context, partials, out=None):
    """Canonical affine sum of extended-Jacobian partial results (device tensor of count*128 bytes)."""
    import torch

    count = partials.numel() * partials.element_size() // 128
    if out is None:
        out = torch.empty(64, dtype=torch.uint8, device=partials.device)
    _lib.check(ctx.handle, ctx.lib.b200zk_g1_sum_dev(ctx.handle, partials.data_ptr(), count, out.data_ptr()))
    return out
