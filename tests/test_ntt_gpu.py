"""GPU parity: b200zk_ntt == fft.Domain.FFT / FFTInverse (gnark-crypto v0.9.1 semantics as restated by the
oracle), bit-exact on the 32*N output bytes, through the C ABI."""
import json
import os

import numpy as np
import pytest

import noir_backend_using_gnark_b200 as zk
from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
VARIANTS = [(inv, dec, cos) for inv in (0, 1) for dec in (zk.DIF, zk.DIT) for cos in (0, 1)]


def run_gpu(ctx, data: np.ndarray, log2n, inverse, dec, coset) -> np.ndarray:
    a = data.copy()
    d = zk.Domain(1 << log2n, ctx)
    (d.FFTInverse if inverse else d.FFT)(a, dec, bool(coset))
    return a


def test_golden_vectors(ctx):
    with open(os.path.join(GOLDEN, "ntt_small.json")) as f:
        for v in json.load(f):
            a = np.frombuffer(bytes.fromhex(v["in"]), dtype=np.uint8)
            got = run_gpu(ctx, a, v["log2n"], v["inverse"], v["decimation"], v["coset"])
            assert got.tobytes().hex() == v["out"], {k: v[k] for k in ("log2n", "inverse", "decimation", "coset")}


@pytest.mark.parametrize("log2n", list(range(0, 15)))
def test_all_variants_small(ctx, log2n):
    a = cref.random_fr(1 << log2n, 0xB2000003 + log2n)
    for inv, dec, cos in VARIANTS:
        want = cref.ntt(a, log2n, inv, dec, cos, nthreads=4)
        got = run_gpu(ctx, a, log2n, inv, dec, cos)
        assert got.tobytes() == want, (log2n, inv, dec, cos)


def test_python_oracle_agrees(ctx):
    # the independent big-int oracle, not just the C port
    log2n = 6
    vals = o.random_fr(1 << log2n, 77)
    a = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    d = o.Domain(1 << log2n)
    for inv, dec, cos in VARIANTS:
        want = (d.fft_inverse if inv else d.fft)(vals, dec, bool(cos))
        got = o.fr_from_mont_bytes(run_gpu(ctx, a, log2n, inv, dec, cos).tobytes())
        assert got == want


@pytest.mark.parametrize("log2n", [16, 17, 19, 20, 21, 22])
def test_multi_pass_sizes(ctx, log2n):
    a = cref.random_fr(1 << log2n, 0xB2000003 + log2n)
    variants = VARIANTS if log2n <= 17 else [(0, zk.DIF, 1), (1, zk.DIT, 1), (1, zk.DIF, 0), (0, zk.DIT, 0)]
    for inv, dec, cos in variants:
        want = cref.ntt(a, log2n, inv, dec, cos, nthreads=cref.ncores())
        got = run_gpu(ctx, a, log2n, inv, dec, cos)
        assert got.tobytes() == want, (log2n, inv, dec, cos)
    cref.load().oracle_domain_release(log2n)


def test_edge_inputs(ctx):
    log2n = 10
    n = 1 << log2n
    zeros = np.zeros(n * 32, dtype=np.uint8)
    for inv, dec, cos in VARIANTS:
        assert not run_gpu(ctx, zeros, log2n, inv, dec, cos).any()
    # delta at 0 -> all ones (Montgomery one), plain forward transform
    delta = zeros.copy()
    one = np.frombuffer(o.fr_to_mont_bytes([1]), dtype=np.uint8)
    delta[:32] = one
    out = run_gpu(ctx, delta, log2n, 0, zk.DIF, 0)
    assert (out.reshape(n, 32) == one).all()
    # maximal element r-1 everywhere
    big = np.tile(np.frombuffer(o.fr_to_mont_bytes([o.R_MOD - 1]), dtype=np.uint8), n)
    assert run_gpu(ctx, big, log2n, 0, zk.DIT, 1).tobytes() == cref.ntt(big, log2n, 0, zk.DIT, 1, 4)


def test_bit_reverse(ctx):
    for log2n in (0, 1, 5, 13):
        a = cref.random_fr(1 << log2n, 5 + log2n)
        got = zk.BitReverse(a.copy(), ctx)
        assert got.tobytes() == cref.bit_reverse(a, log2n)


def test_bad_arguments(ctx):
    lib = zk.load()
    buf = np.zeros(64, dtype=np.uint8)
    assert lib.b200zk_ntt(ctx.handle, buf.ctypes.data, 29, 0, 0, 0) == -3
    assert lib.b200zk_ntt(ctx.handle, None, 4, 0, 0, 0) == -3
    assert lib.b200zk_ntt(ctx.handle, buf.ctypes.data, 1, 0, 7, 0) == -3


def test_device_resident_roundtrip_2_24(ctx):
    """Full-size property check (no oracle needed): FFTInverse(FFT(x)) == x for coset and plain, both orders,
    and the DIF output is the bit-reversal of the DIT-from-bit-reversed output."""
    import torch

    log2n = 24
    n = 1 << log2n
    host = torch.from_numpy(cref.random_fr(n, 0xB2000003))
    x = host.cuda()
    d = zk.Domain(n, ctx)
    for coset in (False, True):
        y = x.clone()
        torch.cuda.synchronize()
        d.FFT(y, zk.DIF, coset)          # natural -> bit-reversed
        d.FFTInverse(y, zk.DIT, coset)   # bit-reversed -> natural
        ctx.sync()
        assert torch.equal(y, x)
        z = x.clone()
        torch.cuda.synchronize()
        d.FFTInverse(z, zk.DIF, coset)
        zk.BitReverse(z, ctx)
        d.FFT(z, zk.DIF, coset)
        zk.BitReverse(z, ctx)
        ctx.sync()
        assert torch.equal(z, x)
    # spot-check against the C oracle on the head of a 2^24 transform would need the full CPU transform (~4 s):
    want = np.frombuffer(cref.ntt(host.numpy(), log2n, 0, zk.DIF, 1, cref.ncores()), dtype=np.uint8)
    y = x.clone()
    torch.cuda.synchronize()
    d.FFT(y, zk.DIF, True)
    ctx.sync()
    assert np.array_equal(y.cpu().numpy(), want)
    cref.load().oracle_domain_release(log2n)


@pytest.mark.parametrize("log2n,world,log2c", [(8, 2, 4), (12, 4, 7), (14, 8, 8), (16, 2, 8), (18, 4, None)])
def test_four_step_halves_emulated_on_one_gpu(ctx, log2n, world, log2c):
    """Every rank's two local halves run on this one GPU; the all-to-all between them is emulated with numpy.
    Checks the gap / block-layout addressing of b200zk_ntt_dist_half_dev for every variant."""
    import torch

    from noir_backend_using_gnark_b200.dist_ntt import ShardLayout

    lay = ShardLayout(log2n, world, log2c)
    lib = zk.load()
    full = cref.random_fr(1 << log2n, 0xB2000003 + log2n)
    chunk = lay.local * 32 // world

    def run_half(src, dst, rank, half, inv, dec, cos):
        rc = lib.b200zk_ntt_dist_half_dev(ctx.handle, src.data_ptr(), dst.data_ptr(), lay.log2n, lay.log2g, rank,
                                          lay.log2c, half, inv, dec, cos)
        assert rc == 0, rc

    variants = VARIANTS if log2n <= 14 else [(0, zk.DIF, 1), (1, zk.DIT, 1), (1, zk.DIF, 0), (0, zk.DIT, 0)]
    for inv, dec, cos in variants:
        xs = [torch.from_numpy(lay.scatter(full, r, column_block=(dec == zk.DIF)).copy()).cuda() for r in range(world)]
        tmps = [torch.empty_like(x) for x in xs]
        torch.cuda.synchronize()
        if dec == zk.DIF:
            for r in range(world):
                run_half(xs[r], xs[r], r, 0, inv, dec, cos)
            ctx.sync()
            for r in range(world):      # all_to_all_single(tmp_r, x_r): chunk j of x_k -> chunk k of tmp_j
                for k in range(world):
                    tmps[r][k * chunk:(k + 1) * chunk] = xs[k][r * chunk:(r + 1) * chunk]
            torch.cuda.synchronize()
            for r in range(world):
                run_half(tmps[r], xs[r], r, 1, inv, dec, cos)
        else:
            for r in range(world):
                run_half(xs[r], tmps[r], r, 0, inv, dec, cos)
            ctx.sync()
            for r in range(world):
                for k in range(world):
                    xs[r][k * chunk:(k + 1) * chunk] = tmps[k][r * chunk:(r + 1) * chunk]
            torch.cuda.synchronize()
            for r in range(world):
                run_half(xs[r], xs[r], r, 1, inv, dec, cos)
        ctx.sync()
        got = lay.gather([x.cpu().numpy() for x in xs], column_block=(dec == zk.DIT))
        want = cref.ntt(full, log2n, inv, dec, cos, cref.ncores())
        assert got.tobytes() == want, (log2n, world, inv, dec, cos)


def test_radix2_and_radix4_kernels_agree(ctx):
    """both pass kernels (radix-4 register kernel and the plain radix-2 one) against the oracle"""
    lib = zk.load()
    for log2n in (5, 11, 15, 19):
        a = cref.random_fr(1 << log2n, 0xB2000003 + log2n)
        for inv, dec, cos in VARIANTS:
            want = cref.ntt(a, log2n, inv, dec, cos, nthreads=cref.ncores())
            for r2 in (1, 0):
                lib.b200zk_ntt_set_radix2(ctx.handle, r2)
                assert run_gpu(ctx, a, log2n, inv, dec, cos).tobytes() == want, (log2n, inv, dec, cos, r2)
    lib.b200zk_ntt_set_radix2(ctx.handle, 0)


def test_large_sizes_roundtrip(ctx):
    """2^25 and 2^26 (4 passes): FFTInverse(FFT(x)) == x, coset, both decimations (size-independent property)."""
    import torch

    for log2n in (25, 26):
        n = 1 << log2n
        x = torch.from_numpy(cref.random_fr(n, 0xB2000003 + log2n)).cuda()
        d = zk.Domain(n, ctx)
        y = x.clone()
        torch.cuda.synchronize()
        d.FFT(y, zk.DIF, True)
        d.FFTInverse(y, zk.DIT, True)
        ctx.sync()
        assert torch.equal(y, x)
        d.FFTInverse(y, zk.DIF, False)
        zk.BitReverse(y, ctx)
        d.FFT(y, zk.DIF, False)
        zk.BitReverse(y, ctx)
        ctx.sync()
        assert torch.equal(y, x)
        del x, y


def test_maximum_size_2_28(ctx):
    """2^28 = the 2-adicity of BN254 fr, the largest domain gnark can build (BASELINE.json config 5).  No oracle at this
    size (8 GiB of input): the transform of a single non-zero coefficient c*X^k must be the geometric sequence
    c*w^(k*j) (sampled against Python big integers), and FFTInverse(FFT(x)) must return a random x, coset included."""
    import torch

    log2n = 28
    n = 1 << log2n
    d = zk.Domain(n, ctx)
    w = pow(o.FR_ROOT_2_28, 1, o.R_MOD)                       # generator of the 2^28 domain itself
    rng = np.random.default_rng(28)
    k, c = int(rng.integers(1, n)), 0x1234567890ABCDEF
    x = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    x[k * 32:(k + 1) * 32] = torch.from_numpy(np.frombuffer(o.fr_to_mont_bytes([c]), dtype=np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    d.FFT(x, zk.DIF, False)                                    # natural in -> bit-reversed out
    ctx.sync()
    for j in [0, 1, 2, n - 1, n // 2] + [int(v) for v in rng.integers(0, n, 40)]:
        pos = int(format(j, "028b")[::-1], 2)
        got = bytes(x[pos * 32:(pos + 1) * 32].cpu().numpy())
        assert got == o.fr_to_mont_bytes([c * pow(w, k * j, o.R_MOD) % o.R_MOD]), j
    d.FFTInverse(x, zk.DIT, False)
    ctx.sync()
    nz = torch.nonzero(x.view(torch.int64)).flatten()
    assert nz.numel() > 0 and int(nz.min()) // 4 == k and int(nz.max()) // 4 == k      # back to c*X^k exactly
    del x
    # random input (valid Montgomery images: top three bits cleared -> < 2^253 < r), coset round trip
    y = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
    y.view(n, 32)[:, 31] &= 0x1F
    ref = y.clone()
    torch.cuda.synchronize()
    d.FFT(y, zk.DIF, True)
    ctx.sync()
    assert not torch.equal(y, ref)
    d.FFTInverse(y, zk.DIT, True)
    ctx.sync()
    assert torch.equal(y, ref)
    del y, ref
    torch.cuda.empty_cache()
