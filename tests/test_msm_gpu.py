"""GPU parity: b200zk_msm_g1 == (*G1Affine).MultiExp (canonical affine result, bit-exact), through the C ABI."""
import json
import os

import numpy as np
import pytest

import noir_backend_using_gnark_b200 as zk
from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gen_point(k: int) -> bytes:
    return o.g1_to_bytes([o.g1_mul(o.G1_GEN, k)])


def structured(n: int, a: int = 0xB2000002 % 9973, b: int = 7919) -> np.ndarray:
    """P_i = (a + i*b)*G (SURVEY.md §8c self-checking bases)."""
    return cref.g1_arith_progression(gen_point(a), gen_point(b), n)


def closed_form(scalars_mont: np.ndarray, n: int, a: int = 0xB2000002 % 9973, b: int = 7919) -> bytes:
    s = o.fr_from_mont_bytes(scalars_mont[: n * 32].tobytes())
    k = sum(si * (a + i * b) for i, si in enumerate(s)) % o.R_MOD
    return o.g1_to_bytes([o.g1_mul(o.G1_GEN, k)])


def test_golden_vectors(ctx):
    with open(os.path.join(GOLDEN, "msm_small.json")) as f:
        for v in json.load(f):
            srs = zk.SRS(bytes.fromhex(v["points"]), ctx)
            got = zk.MultiExp(srs, bytes.fromhex(v["scalars"]))
            assert got.hex() == v["out"], v["n"]
            srs.close()


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 1000, 4097])
def test_small_sizes_vs_oracle(ctx, n):
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + n)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=4)
    assert zk.MultiExp(srs, sc) == want
    assert want == closed_form(sc, n)
    srs.close()


def test_empty_and_prefix(ctx):
    pts = structured(64)
    sc = cref.random_fr(64, 3)
    srs = zk.SRS(pts, ctx)
    assert zk.MultiExp(srs, sc[:0], n=0) == b"\0" * 64           # empty sum = infinity = (0,0)
    assert zk.MultiExp(srs, sc, n=10) == cref.msm(pts, sc, 10)   # kzg.Commit uses a prefix of the SRS
    lib = zk.load()
    out = np.zeros(64, dtype=np.uint8)
    assert lib.b200zk_msm_g1(ctx.handle, srs.handle, sc.ctypes.data, 65, out.ctypes.data) == -3  # n > len(bases)
    srs.close()


def test_special_scalars_and_points(ctx):
    n = 600
    pts = structured(n).copy()
    pts[5 * 64 : 6 * 64] = 0                       # infinity among the bases
    pts[9 * 64 : 10 * 64] = pts[8 * 64 : 9 * 64]   # duplicate point
    vals = o.random_fr(n, 17)
    vals[0] = 0
    vals[1] = 1
    vals[2] = o.R_MOD - 1
    vals[3] = (1 << 253) % o.R_MOD
    vals[8] = vals[9]                              # same point, same scalar: exercises the doubling path
    vals[20] = 1 << 15                             # digit exactly 2^(c-1) for c = 16
    sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=4)
    for c in (0, 6, 9, 13, 16):
        zk.load().b200zk_msm_set_window(ctx.handle, c)
        assert zk.MultiExp(srs, sc) == want, c
    zk.load().b200zk_msm_set_window(ctx.handle, 0)
    # all-zero scalars -> infinity
    assert zk.MultiExp(srs, np.zeros(n * 32, dtype=np.uint8)) == b"\0" * 64
    srs.close()


def test_cancellation_to_infinity(ctx):
    # s*P + (r-s)*P = infinity
    n = 2
    p = gen_point(12345)
    pts = np.frombuffer(p + p, dtype=np.uint8)
    s = 0x1234567890ABCDEF1234567890ABCDEF
    sc = np.frombuffer(o.fr_to_mont_bytes([s, o.R_MOD - s]), dtype=np.uint8)
    srs = zk.SRS(pts, ctx)
    assert zk.MultiExp(srs, sc) == b"\0" * 64
    srs.close()


@pytest.mark.parametrize("kind", ["all_equal", "witness_like", "same_point"])
def test_skewed_inputs(ctx, kind):
    """Bucket skew (SURVEY.md §8d adversarial sets): long runs go through the cooperative path."""
    n = 1 << 15
    pts = structured(n)
    if kind == "all_equal":
        v = o.random_fr(1, 5)[0]
        sc = np.tile(np.frombuffer(o.fr_to_mont_bytes([v]), dtype=np.uint8), n)
    elif kind == "witness_like":
        rnd = o.random_fr(n // 4, 6)
        vals = [0] * (n // 2) + [x % (1 << 16) for x in rnd] + rnd
        sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    else:
        pts = np.tile(np.frombuffer(gen_point(777), dtype=np.uint8), n)
        sc = cref.random_fr(n, 8)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    for c in (0, 8):
        zk.load().b200zk_msm_set_window(ctx.handle, c)
        assert zk.MultiExp(srs, sc) == want, (kind, c)
    zk.load().b200zk_msm_set_window(ctx.handle, 0)
    srs.close()


@pytest.mark.parametrize("log2n", [16, 18])
def test_vs_c_oracle_medium(ctx, log2n):
    n = 1 << log2n
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    assert zk.MultiExp(srs, sc) == want
    srs.close()


@pytest.mark.parametrize("log2n", [20, 22])
def test_closed_form_large(ctx, log2n):
    """Size-independent check: with P_i = (a+i*b)G the MSM must equal ((sum s_i (a+i*b)) mod r) * G."""
    import torch

    n = 1 << log2n
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + log2n)
    srs = zk.SRS(pts, ctx)
    want = closed_form(sc, n)
    assert zk.MultiExp(srs, sc) == want
    # device-resident + sharded: two half-range partials combined == single result
    d_sc = torch.from_numpy(sc).cuda()
    torch.cuda.synchronize()
    half = n // 2
    parts = torch.empty(256, dtype=torch.uint8, device="cuda")
    zk.MultiExp(srs, d_sc[: half * 32], n=half, first_base=0, out=parts[:128], partial=True)
    zk.MultiExp(srs, d_sc[half * 32 :], n=half, first_base=half, out=parts[128:], partial=True)
    res = zk.SumPartials(ctx, parts)
    ctx.sync()
    assert res.cpu().numpy().tobytes() == want
    srs.close()


def test_srs_generate_matches_oracle(ctx):
    """kzg.NewSRS on the device: G1[i] = alpha^i * G; and Commit(p) == p(alpha) * G."""
    alpha = o.random_fr(1, 0xB2000005)[0]
    amont = o.fr_to_mont_bytes([alpha])
    n = 70
    srs = zk.SRS.NewSRS(n, amont, ctx)
    got = o.g1_from_bytes(srs.download())
    want = [o.g1_mul(o.G1_GEN, pow(alpha, i, o.R_MOD)) for i in range(n)]
    assert got == want
    shard = zk.SRS.NewSRS(10, amont, ctx, first=60)
    assert o.g1_from_bytes(shard.download()) == want[60:70]
    coeffs = o.random_fr(n, 4)
    com = zk.Commit(np.frombuffer(o.fr_to_mont_bytes(coeffs), dtype=np.uint8), srs)
    p_alpha = sum(c * pow(alpha, i, o.R_MOD) for i, c in enumerate(coeffs)) % o.R_MOD
    assert com == o.g1_to_bytes([o.g1_mul(o.G1_GEN, p_alpha)])
    srs.close()
    shard.close()


@pytest.mark.parametrize("log2n,c", [(10, 0), (12, 10), (14, 0), (16, 13), (18, 0)])
def test_precomputed_window_table(ctx, log2n, c):
    """b200zk_bases_precompute: same canonical result through the single-bucket-set path (prefixes, shards, skew)."""
    n = 1 << log2n
    pts = structured(n).copy()
    pts[7 * 64: 8 * 64] = 0                      # a point at infinity among the bases
    sc = cref.random_fr(n, 0xB2000001 + log2n)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    assert zk.MultiExp(srs, sc) == want          # classic path
    srs.precompute(c)
    assert zk.MultiExp(srs, sc) == want          # table path
    m = n - 5
    assert zk.MultiExp(srs, sc, n=m) == cref.msm(pts, sc, m, nthreads=cref.ncores())   # kzg.Commit prefix
    assert zk.MultiExp(srs, sc, n=3) == cref.msm(pts, sc, 3)                            # small n falls back to classic
    if log2n >= 12:
        import torch

        d_sc = torch.from_numpy(sc).cuda()
        torch.cuda.synchronize()
        half = n // 2
        parts = torch.empty(256, dtype=torch.uint8, device="cuda")
        zk.MultiExp(srs, d_sc[: half * 32], n=half, first_base=0, out=parts[:128], partial=True)
        zk.MultiExp(srs, d_sc[half * 32:], n=half, first_base=half, out=parts[128:], partial=True)
        res = zk.SumPartials(ctx, parts)
        ctx.sync()
        assert res.cpu().numpy().tobytes() == want
    # skewed scalars through the table path
    v = o.random_fr(1, 5)[0]
    eq = np.tile(np.frombuffer(o.fr_to_mont_bytes([v]), dtype=np.uint8), n)
    assert zk.MultiExp(srs, eq) == cref.msm(pts, eq, n, nthreads=cref.ncores())
    srs.close()


def test_full_size_2_24_table_vs_classic_vs_shards(ctx):
    """BASELINE size (2^24 points): no oracle at this size — the window-table path, the classic path and the
    two-shard combination must agree byte for byte (three different summation orders of the same MSM)."""
    import torch

    n = 1 << 24
    alpha = o.fr_to_mont_bytes([o.random_fr(1, 0xB2000005)[0]])
    srs = zk.SRS.NewSRS(n, alpha, ctx)
    sc = torch.from_numpy(cref.random_fr(n, 0xB2000001)).cuda()
    torch.cuda.synchronize()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    zk.MultiExp(srs, sc, n=n, out=out)
    ctx.sync()
    classic = out.cpu().numpy().tobytes()
    srs.precompute()
    zk.MultiExp(srs, sc, n=n, out=out)
    ctx.sync()
    table = out.cpu().numpy().tobytes()
    half = n // 2
    parts = torch.empty(256, dtype=torch.uint8, device="cuda")
    zk.MultiExp(srs, sc[: half * 32], n=half, first_base=0, out=parts[:128], partial=True)
    zk.MultiExp(srs, sc[half * 32:], n=half, first_base=half, out=parts[128:], partial=True)
    res = zk.SumPartials(ctx, parts)
    ctx.sync()
    assert classic == table == res.cpu().numpy().tobytes()
    assert classic != b"\0" * 64
    srs.close()


def test_srs_compressed_roundtrip_matches_gnark_encoding(ctx):
    """kzg.SRS file format (compressed G1): device compression == the oracle's G1Affine.Bytes(), device decompression
    (fp square root + sign flag) gives the points back; off-curve input is rejected."""
    from oracle import plonk as pl

    alpha = o.random_fr(1, 0xB2000005)[0]
    n = 300
    srs = zk.SRS.NewSRS(n, o.fr_to_mont_bytes([alpha]), ctx)
    pts = o.g1_from_bytes(srs.download())
    comp = srs.download_compressed()
    assert comp == b"".join(pl.g1_compress(p) for p in pts)
    comp = bytearray(comp)
    comp[5 * 32: 6 * 32] = pl.g1_compress(None)                 # an infinity entry
    back = zk.SRS.FromCompressed(bytes(comp), ctx)
    want = list(pts)
    want[5] = None
    assert o.g1_from_bytes(back.download()) == want
    # x = 4 is not on the curve (4^3 + 3 = 67 is a non-residue mod p? checked below) -> rejected
    x = 4
    while pow((x ** 3 + 3) % o.P_MOD, (o.P_MOD - 1) // 2, o.P_MOD) == 1:
        x += 1
    badpt = bytearray(x.to_bytes(32, "big"))
    badpt[0] |= 0x80
    with pytest.raises(zk.B200zkError):
        zk.SRS.FromCompressed(bytes(badpt), ctx)
    srs.close()
    back.close()


def test_scatter_with_zero_digits_and_long_runs(ctx):
    """the counting-sort front end on skewed scalars (zero digits, one very long run), classic windows and table"""
    n = 1 << 17
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + 99)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    srs = zk.SRS(pts, ctx)
    for table in (False, True):
        if table:
            srs.precompute()
        for flat in (0,):
            assert zk.MultiExp(srs, sc) == want, (table, flat)
            # skewed scalars: half of them zero (zero digits -> the sentinel bucket of the sort-based front end), a
            # quarter all-equal (one very long run)
            sk = sc.copy().reshape(n, 32)
            sk[::2] = 0
            sk[1::4] = sk[1]
            sk = sk.reshape(-1)
            assert zk.MultiExp(srs, sk) == cref.msm(pts, sk, n, nthreads=cref.ncores()), (table, flat, "skewed")
    srs.close()


@pytest.mark.parametrize("chunks", [2, 3, 4, 8])
def test_host_scalar_msm_split_for_copy_overlap(ctx, chunks):
    """b200zk_msm_g1 splits host scalars by point range (copy of chunk i+1 under the MSM of chunk i): the result must
    not depend on the split — classic windows, window table, ragged last chunk, pageable and pinned host buffers."""
    import torch

    lib = zk.load()
    n = (1 << 16) + 777
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + 5)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    srs = zk.SRS(pts, ctx)
    try:
        for table in (False, True):
            if table:
                srs.precompute()
            lib.b200zk_msm_set_host_chunks(ctx.handle, 1)
            assert zk.MultiExp(srs, sc) == want
            lib.b200zk_msm_set_host_chunks(ctx.handle, chunks)
            assert zk.MultiExp(srs, sc) == want, (table, chunks)
            pinned = torch.from_numpy(sc.copy()).pin_memory()
            assert zk.MultiExp(srs, pinned.numpy()) == want
            assert zk.MultiExp(srs, sc[: 5000 * 32], n=5000) == cref.msm(pts, sc, 5000, nthreads=2)   # too small: not split
    finally:
        lib.b200zk_msm_set_host_chunks(ctx.handle, 0)
        srs.close()


def test_maximum_sweep_size_2_26(ctx):
    """2^26 points (the top of BASELINE.json's MSM sweep): no oracle at this size — classic windows, the window table
    and the sum of four point-range shards are three different summation orders and must agree byte for byte; with all
    scalars equal to s the result must also be the closed form s*(alpha^n - 1)/(alpha - 1)*G."""
    import torch

    n = 1 << 26
    alpha_int = o.random_fr(1, 0xB2000005)[0]
    srs = zk.SRS.NewSRS(n, o.fr_to_mont_bytes([alpha_int]), ctx)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
    sc.view(n, 32)[:, 31] &= 0x1F                                # < 2^253 < r: valid Montgomery images
    torch.cuda.synchronize()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    zk.MultiExp(srs, sc, n=n, out=out)
    ctx.sync()
    classic = out.cpu().numpy().tobytes()
    srs.precompute()
    zk.MultiExp(srs, sc, n=n, out=out)
    ctx.sync()
    table = out.cpu().numpy().tobytes()
    q = n // 4
    parts = torch.empty(4 * 128, dtype=torch.uint8, device="cuda")
    for i in range(4):
        zk.MultiExp(srs, sc[i * q * 32:(i + 1) * q * 32], n=q, first_base=i * q, out=parts[128 * i:128 * (i + 1)], partial=True)
    res = zk.SumPartials(ctx, parts)
    ctx.sync()
    assert classic == table == res.cpu().numpy().tobytes() and classic != b"\0" * 64
    s = 0x0123456789ABCDEF0123456789ABCDEF % o.R_MOD
    eq = torch.from_numpy(np.frombuffer(o.fr_to_mont_bytes([s]), dtype=np.uint8).copy()).cuda().repeat(n)
    torch.cuda.synchronize()
    zk.MultiExp(srs, eq, n=n, out=out)
    ctx.sync()
    geo = (pow(alpha_int, n, o.R_MOD) - 1) * pow(alpha_int - 1, -1, o.R_MOD) % o.R_MOD
    assert out.cpu().numpy().tobytes() == o.g1_to_bytes([o.g1_mul(o.G1_GEN, s * geo % o.R_MOD)])
    srs.close()
    del sc, eq, parts
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", [1, 2, 5, 33, 100, 600, 1000, 4097, 6000])
def test_small_msm_path_with_window_table(ctx, n):
    """With a window table, MSMs of the size of the reference's own test circuits take a two-launch path (one thread
    per (point, window) term).  Same result as the oracle, as the bucket pipeline on the same table and as classic
    windows — including infinity bases, duplicate points, 0 / 1 / -1 scalars, cancellation and sub-ranges."""
    lib = zk.load()
    pts = structured(n).copy()
    vals = o.random_fr(n, 0xB2000001 + n)
    if n >= 33:
        pts[5 * 64: 6 * 64] = 0
        pts[9 * 64: 10 * 64] = pts[8 * 64: 9 * 64]
        vals[0], vals[1], vals[2], vals[3] = 0, 1, o.R_MOD - 1, (1 << 253) % o.R_MOD
        vals[8] = vals[9]
        pts[12 * 64: 13 * 64] = pts[11 * 64: 12 * 64]
        vals[12] = o.R_MOD - vals[11]                                  # cancels point 11
    sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    want = cref.msm(pts, sc, n, nthreads=4)
    srs = zk.SRS(pts, ctx)
    try:
        assert zk.MultiExp(srs, sc) == want                             # classic windows
        srs.precompute()
        assert zk.MultiExp(srs, sc) == want                             # small path
        lib.b200zk_msm_set_small_path(ctx.handle, 0)
        assert zk.MultiExp(srs, sc) == want                             # bucket pipeline on the same table
        lib.b200zk_msm_set_small_path(ctx.handle, 1)
        assert zk.MultiExp(srs, np.zeros(n * 32, dtype=np.uint8)) == b"\0" * 64
        if n >= 100:
            import torch

            d = torch.from_numpy(sc.copy()).cuda()
            out = torch.zeros(64, dtype=torch.uint8, device="cuda")
            zk.MultiExp(srs, d[40 * 32: 90 * 32], n=50, first_base=40, out=out)
            ctx.sync()
            assert out.cpu().numpy().tobytes() == cref.msm(pts[40 * 64: 90 * 64], sc[40 * 32: 90 * 32], 50, nthreads=2)
    finally:
        lib.b200zk_msm_set_small_path(ctx.handle, 1)
        srs.close()


@pytest.mark.parametrize("log2n", [13, 17])
def test_bucket_reduction_chunk_sizes_agree(ctx, log2n):
    """the bucket reduction with 8-bucket and 32-bucket running-sum chunks (latency- vs throughput-oriented), classic
    windows and window table, against the oracle"""
    lib = zk.load()
    n = (1 << log2n) + 5
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + 31 + log2n)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    srs = zk.SRS(pts, ctx)
    try:
        for table in (False, True):
            if table:
                srs.precompute()
            for cl in (3, 4, 5, 0):
                lib.b200zk_msm_set_reduce_chunk(ctx.handle, cl)
                assert zk.MultiExp(srs, sc) == want, (table, cl)
    finally:
        lib.b200zk_msm_set_reduce_chunk(ctx.handle, 0)
        srs.close()


def test_pseudo_random_bases(ctx):
    """SURVEY.md §8d bases (ii): points by try-and-increment (no structure between the bases), seed 0xB2000002, against
    the C oracle — uniform and witness-like scalars, classic windows and window table."""
    n = 1 << 13
    pts = np.frombuffer(o.g1_to_bytes(o.g1_pseudo_random_points(n, 0xB2000002)), dtype=np.uint8)
    uniform = cref.random_fr(n, 0xB2000001)
    vals = o.random_fr(n, 0xB2000001 + 1)
    for i in range(n):
        if i % 4 < 2:
            vals[i] = 0
        elif i % 4 == 2:
            vals[i] &= 0xFFFF
    witness_like = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    srs = zk.SRS(pts, ctx)
    try:
        for table in (False, True):
            if table:
                srs.precompute()
            for sc in (uniform, witness_like):
                assert zk.MultiExp(srs, sc) == cref.msm(pts, sc, n, nthreads=cref.ncores())
    finally:
        srs.close()


def test_small_msm_on_the_head_of_a_large_table(ctx):
    """the reference's default: a large SRS (here 2^15 points) serving tiny commitments — the head of the bases has a
    narrow-window table of its own; results equal the oracle for sizes on both sides of its 2048-point limit and for a
    sub-range that straddles it"""
    import torch

    n = 1 << 15
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001 + 77)
    srs = zk.SRS(pts, ctx).precompute()
    try:
        for m in (1, 11, 67, 2048, 2049, 5000):
            assert zk.MultiExp(srs, sc[: m * 32], n=m) == cref.msm(pts, sc, m, nthreads=4), m
        d = torch.from_numpy(sc.copy()).cuda()
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        for first, m in ((100, 50), (2000, 100), (3000, 10)):
            zk.MultiExp(srs, d[first * 32:(first + m) * 32], n=m, first_base=first, out=out)
            ctx.sync()
            want = cref.msm(pts[first * 64:(first + m) * 64], sc[first * 32:(first + m) * 32], m, nthreads=2)
            assert out.cpu().numpy().tobytes() == want, (first, m)
    finally:
        srs.close()


# ---- pair rounds: batched affine pre-summation of the bucket runs (b200zk_msm_set_pair_rounds) ----------------------
@pytest.fixture
def pair_rounds(ctx):
    lib = zk.load()

    def set_rounds(r):
        assert lib.b200zk_msm_set_pair_rounds(ctx.handle, r) == 0

    yield set_rounds
    lib.b200zk_msm_set_pair_rounds(ctx.handle, -1)
    lib.b200zk_msm_set_small_path(ctx.handle, 1)
    lib.b200zk_msm_set_window(ctx.handle, 0)
    lib.b200zk_msm_set_host_chunks(ctx.handle, 0)


@pytest.mark.parametrize("rounds", [1, 2, 3, 5])
def test_pair_rounds_special_points_and_scalars(ctx, pair_rounds, rounds):
    """Infinity among the bases, duplicate points with equal scalars (the tangent case of the affine addition),
    P + (-P) inside a run, zero / one / r-1 scalars, every window size, with and without the window table."""
    lib = zk.load()
    n = 600
    pts = structured(n).copy()
    pts[5 * 64 : 6 * 64] = 0
    pts[9 * 64 : 10 * 64] = pts[8 * 64 : 9 * 64]
    pts[11 * 64 : 12 * 64] = pts[8 * 64 : 9 * 64]
    vals = o.random_fr(n, 17)
    vals[0] = 0
    vals[1] = 1
    vals[2] = o.R_MOD - 1
    vals[3] = (1 << 253) % o.R_MOD
    vals[8] = vals[9]
    vals[11] = o.R_MOD - vals[8]                    # -P next to P, P in the same buckets
    vals[20] = 1 << 15
    sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=4)
    pair_rounds(rounds)
    lib.b200zk_msm_set_small_path(ctx.handle, 0)
    for table in (False, True):
        if table:
            srs.precompute()
        for c in (0, 6, 9, 13, 16):
            lib.b200zk_msm_set_window(ctx.handle, c)
            assert zk.MultiExp(srs, sc) == want, (table, c)
        lib.b200zk_msm_set_window(ctx.handle, 0)
        assert zk.MultiExp(srs, np.zeros(n * 32, dtype=np.uint8)) == b"\0" * 64
    srs.close()


@pytest.mark.parametrize("rounds", [1, 3, 6])
@pytest.mark.parametrize("kind", ["all_equal", "witness_like", "same_point"])
def test_pair_rounds_skewed_inputs(ctx, pair_rounds, rounds, kind):
    """Long runs after the pair rounds still go through the cooperative path (run bounds shifted by the rounds)."""
    n = 1 << 15
    pts = structured(n)
    if kind == "all_equal":
        v = o.random_fr(1, 5)[0]
        sc = np.tile(np.frombuffer(o.fr_to_mont_bytes([v]), dtype=np.uint8), n)
    elif kind == "witness_like":
        rnd = o.random_fr(n // 4, 6)
        vals = [0] * (n // 2) + [x % (1 << 16) for x in rnd] + rnd
        sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
    else:
        pts = np.tile(np.frombuffer(gen_point(777), dtype=np.uint8), n)   # every pair is a doubling
        sc = cref.random_fr(n, 8)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    pair_rounds(rounds)
    for c in (0, 8):
        zk.load().b200zk_msm_set_window(ctx.handle, c)
        assert zk.MultiExp(srs, sc) == want, (kind, c)
    zk.load().b200zk_msm_set_window(ctx.handle, 0)
    srs.precompute()
    assert zk.MultiExp(srs, sc) == want, (kind, "table")
    srs.close()


@pytest.mark.parametrize("log2n", [16, 18])
def test_pair_rounds_vs_c_oracle_medium(ctx, pair_rounds, log2n):
    lib = zk.load()
    n = (1 << log2n) + 13
    pts = structured(n)
    sc = cref.random_fr(n, 0xB2000001)
    srs = zk.SRS(pts, ctx)
    want = cref.msm(pts, sc, n, nthreads=cref.ncores())
    for table in (False, True):
        if table:
            srs.precompute()
        for rounds in (0, 1, 2, 3, 4):
            pair_rounds(rounds)
            assert zk.MultiExp(srs, sc) == want, (table, rounds)
        pair_rounds(2)
        lib.b200zk_msm_set_host_chunks(ctx.handle, 3)   # host-scalar chunks resume from the buckets of the chunk before
        assert zk.MultiExp(srs, sc) == want, (table, "chunks")
        lib.b200zk_msm_set_host_chunks(ctx.handle, 0)
    srs.close()


def test_pair_rounds_full_size_2_24(ctx, pair_rounds):
    """2^24 points (the bench configuration): pair rounds off / automatic / forced agree byte for byte, and with all
    scalars equal the result is the closed form s*(alpha^n - 1)/(alpha - 1)*G."""
    import torch

    n = 1 << 24
    alpha_int = o.random_fr(1, 0xB2000005)[0]
    srs = zk.SRS.NewSRS(n, o.fr_to_mont_bytes([alpha_int]), ctx)
    srs.precompute()
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda")
    sc.view(n, 32)[:, 31] &= 0x1F
    torch.cuda.synchronize()
    got = {}
    for rounds in (0, -1, 1, 3, 4):
        pair_rounds(rounds)
        res = zk.MultiExp(srs, sc, n=n)
        ctx.sync()
        got[rounds] = bytes(res.cpu().numpy())
    assert len(set(got.values())) == 1, {k: v.hex()[:16] for k, v in got.items()}
    s = o.random_fr(1, 99)[0]
    one = np.frombuffer(o.fr_to_mont_bytes([s]), dtype=np.uint8)
    sc.view(n, 32)[:] = torch.from_numpy(one.copy()).cuda()
    torch.cuda.synchronize()
    geo = (pow(alpha_int, n, o.R_MOD) - 1) * pow(alpha_int - 1, -1, o.R_MOD) % o.R_MOD
    want = o.g1_to_bytes([o.g1_mul(o.G1_GEN, s * geo % o.R_MOD)])
    pair_rounds(3)
    res = zk.MultiExp(srs, sc, n=n)
    ctx.sync()
    assert bytes(res.cpu().numpy()) == want
    srs.close()


def test_two_product_sweep_self_check(ctx):
    """b200zk_microbench(ctx, 7): every thread compares fe_mul2add(a, b, c, d) with fe_add(fe_mul(a, b), fe_mul(c, d)) on
    lane-dependent and on near-modulus operands and traps on a mismatch (a trap surfaces as a CUDA error here); the rate it
    returns must beat the plain product's (one reduction for two products)."""
    two = ctx.microbench(7)
    one = ctx.microbench(5)
    assert two > 1.15 * one, (two, one)


# ---- bucket-range scatter passes (b200zk_msm_set_scatter_passes) -----------------------------------------------------
@pytest.mark.parametrize("passes", [2, 7, 64])
def test_scatter_passes_small_and_skewed(ctx, passes):
    """The scatter in bucket-range passes (key records parked by the histogram kernel) must place exactly the entries of
    the single-pass scatter: special scalars and points, every window size (classic windows use the warp-aggregated top
    window), the window table, all-equal scalars (one bucket per window takes everything), host-scalar chunks, and on
    top of pair rounds."""
    lib = zk.load()
    try:
        n = 3000
        pts = structured(n).copy()
        pts[5 * 64 : 6 * 64] = 0
        vals = o.random_fr(n, 23)
        vals[0] = 0
        vals[1] = 1
        vals[2] = o.R_MOD - 1
        vals[7] = 1 << 15
        sc = np.frombuffer(o.fr_to_mont_bytes(vals), dtype=np.uint8)
        same = np.tile(sc[32 * 11 : 32 * 12], n)
        srs = zk.SRS(pts, ctx)
        want, want_same = cref.msm(pts, sc, n, nthreads=4), cref.msm(pts, same, n, nthreads=4)
        lib.b200zk_msm_set_small_path(ctx.handle, 0)
        for table in (False, True):
            if table:
                srs.precompute()
            for c in (0, 6, 11, 16):
                lib.b200zk_msm_set_window(ctx.handle, c)
                lib.b200zk_msm_set_scatter_passes(ctx.handle, 1)
                assert zk.MultiExp(srs, sc) == want, (table, c, "single")
                lib.b200zk_msm_set_scatter_passes(ctx.handle, passes)
                assert zk.MultiExp(srs, sc) == want, (table, c)
                assert zk.MultiExp(srs, same) == want_same, (table, c, "all equal")
            lib.b200zk_msm_set_window(ctx.handle, 0)
            lib.b200zk_msm_set_pair_rounds(ctx.handle, 2)
            assert zk.MultiExp(srs, sc) == want, (table, "pair rounds")
            lib.b200zk_msm_set_pair_rounds(ctx.handle, -1)
        srs.close()
        n = (1 << 16) + 5
        pts = structured(n)
        sc = cref.random_fr(n, 0xB2000001 + 9)
        srs = zk.SRS(pts, ctx)
        srs.precompute()
        want = cref.msm(pts, sc, n, nthreads=cref.ncores())
        lib.b200zk_msm_set_host_chunks(ctx.handle, 3)
        assert zk.MultiExp(srs, sc) == want
        srs.close()
    finally:
        lib.b200zk_msm_set_scatter_passes(ctx.handle, 0)
        lib.b200zk_msm_set_small_path(ctx.handle, 1)
        lib.b200zk_msm_set_window(ctx.handle, 0)
        lib.b200zk_msm_set_host_chunks(ctx.handle, 0)
        lib.b200zk_msm_set_pair_rounds(ctx.handle, -1)
