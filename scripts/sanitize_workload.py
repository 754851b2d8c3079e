"""Workload for compute-sanitizer (profiles/r02_sanitizer.md): small instances of every kernel family.
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_workload.py [parts]
parts (default all): ntt msm prove ffi distntt distprove   (the last two emulate 2 ranks on one GPU; distprove needs the
ranks' host threads to run concurrently, which memcheck allows and racecheck / synccheck serialise — see the .md)"""
import os, sys, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ["CUDA_MODULE_LOADING"] = "EAGER"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from noir_backend_using_gnark_b200.dist_ntt import ShardLayout

parts = sys.argv[1:] or ["ntt", "msm", "prove", "distntt", "distprove"]
rng = np.random.default_rng(7)


def fr(n):
    limbs = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    limbs[:, 3] &= (1 << 60) - 1
    return limbs.view(np.uint8).reshape(-1).copy()


ctx = zk.Context(0)
lib = zk.load()
alpha = zkp.fr_to_mont([0x1234567890ABCDEF])
if "ntt" in parts:
    for lg in (1, 5, 11, 14):
        d = zk.Domain(1 << lg, ctx)
        a = fr(1 << lg)
        for inv in (0, 1):
            for dec in (zk.DIF, zk.DIT):
                for cos in (False, True):
                    x = a.copy()
                    (d.FFTInverse if inv else d.FFT)(x, dec, cos)
        zk.BitReverse(a.copy(), ctx)
    lib.b200zk_ntt_set_radix2(ctx.handle, 1)
    zk.Domain(1 << 12, ctx).FFT(fr(1 << 12), zk.DIF, True)
    lib.b200zk_ntt_set_radix2(ctx.handle, 0)
    print("ntt ok", flush=True)
if "msm" in parts:
    n = 1 << 14
    srs = zk.SRS.NewSRS(n, alpha, ctx)
    sc = fr(n)
    r0 = zk.MultiExp(srs, sc)                      # classic windows
    same = np.tile(sc[:32], n)                     # all-equal scalars: the long-run path
    r1 = zk.MultiExp(srs, same)
    srs.precompute()
    assert zk.MultiExp(srs, sc) == r0              # window table (bucket pipeline)
    assert zk.MultiExp(srs, same) == r1
    lib.b200zk_msm_set_small_path(ctx.handle, 1)
    zk.MultiExp(srs, sc[: 100 * 32], n=100)        # tiny path
    lib.b200zk_msm_set_host_chunks(ctx.handle, 3)
    assert zk.MultiExp(srs, sc) == r0              # chunked host path (copy stream + partial sums)
    lib.b200zk_msm_set_host_chunks(ctx.handle, 0)
    for rounds in (1, 3):                          # batched-affine pair rounds: padded runs, pair kernels, shifted run bounds
        lib.b200zk_msm_set_pair_rounds(ctx.handle, rounds)
        lib.b200zk_msm_set_small_path(ctx.handle, 0)
        assert zk.MultiExp(srs, sc) == r0
        assert zk.MultiExp(srs, same) == r1
        lib.b200zk_msm_set_window(ctx.handle, 9)   # classic windows
        assert zk.MultiExp(srs, sc) == r0
        lib.b200zk_msm_set_window(ctx.handle, 0)
    lib.b200zk_msm_set_pair_rounds(ctx.handle, -1)
    for passes in (2, 9):                          # scatter in bucket-range passes over the parked key records
        lib.b200zk_msm_set_scatter_passes(ctx.handle, passes)
        assert zk.MultiExp(srs, sc) == r0
        assert zk.MultiExp(srs, same) == r1
        lib.b200zk_msm_set_window(ctx.handle, 9)
        assert zk.MultiExp(srs, sc) == r0
        lib.b200zk_msm_set_window(ctx.handle, 0)
    lib.b200zk_msm_set_scatter_passes(ctx.handle, 0)
    lib.b200zk_msm_set_small_path(ctx.handle, 1)
    comp = srs.download_compressed(0, 64)
    zk.SRS.FromCompressed(comp, ctx).close()
    srs.close()
    print("msm ok", flush=True)


def chain(gates, nb_public=2):
    """x_{i+1} = x_i^2 + x_i + c_i"""
    R = zkp.R_MOD
    x = [int(v) for v in rng.integers(1, 1 << 62, size=nb_public)]
    ql, qr, qm, qo, qk, a, b, c = [], [], [], [], [], [], [], []
    cur = nb_public - 1
    for i in range(gates):
        ci = int(rng.integers(1, 1 << 62))
        x.append((x[cur] * x[cur] + x[cur] + ci) % R)
        ql.append(1); qr.append(0); qm.append(1); qo.append(R - 1); qk.append(ci)
        a.append(cur); b.append(cur); c.append(len(x) - 1)
        cur = len(x) - 1
    return zkp.SparseR1CS(nb_public, len(x) - nb_public, ql, qr, qm, qo, qk, a, b, c), x


blind = fr(9)
if "prove" in parts:
    cs, x = chain(200)
    srs = zk.SRS.NewSRS(259, alpha, ctx).precompute()
    pk = zkp.ProvingKey.Setup(cs, srs, ctx)
    p1 = pk.Prove(zkp.fr_to_mont(x), blind)                    # 3 commitment lanes
    lib.b200zk_plonk_set_commit_lanes(ctx.handle, 1)
    assert pk.Prove(zkp.fr_to_mont(x), blind).blob == p1.blob
    lib.b200zk_plonk_set_commit_lanes(ctx.handle, 3)
    lib.b200zk_msm_set_small_path(ctx.handle, 0)               # force the bucket pipeline on the lanes
    assert pk.Prove(zkp.fr_to_mont(x), blind).blob == p1.blob
    lib.b200zk_msm_set_small_path(ctx.handle, 1)
    pk.close(); srs.close()
    print("prove ok", flush=True)
if "distntt" in parts:
    import torch
    lay = ShardLayout(14, 2, 8)
    full = fr(1 << 14)
    for inv, dec in ((0, zk.DIF), (1, zk.DIT)):
        xs = [torch.from_numpy(lay.scatter(full, r, column_block=(dec == zk.DIF)).copy()).cuda() for r in range(2)]
        bufs = [torch.empty_like(x) for x in xs]
        torch.cuda.synchronize()
        ptrs = (C.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
        for r in range(2):   # fused exchange: rank r's last pass stores into both "peers'" buffers
            assert lib.b200zk_ntt_dist_half0_p2p_dev(ctx.handle, xs[r].data_ptr(), ptrs, 14, 1, r, 8, inv, dec, 1) == 0
        ctx.sync()
        for r in range(2):
            dst = xs[r] if dec == zk.DIF else bufs[r]
            assert lib.b200zk_ntt_dist_half_dev(ctx.handle, bufs[r].data_ptr(), dst.data_ptr(), 14, 1, r, 8, 1, inv, dec, 1) == 0
        ctx.sync()
    print("distntt ok", flush=True)
if "distprove" in parts:
    cs, x = chain(3000)
    ctxs = [ctx, zk.Context(0)]
    srss = [zk.SRS.NewSRS(4099, alpha, c).precompute() for c in ctxs]
    pks = [zkp.ProvingKey.Setup(cs, s, c) for s, c in zip(srss, ctxs)]
    single = pks[0].Prove(zkp.fr_to_mont(x), blind).blob
    zkp.ProvingKey.JoinLocal(pks)
    res = [None, None]

    def run(k):
        res[k] = pks[k].Prove(zkp.fr_to_mont(x) if k == 0 else None, blind if k == 0 else None).blob

    th = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert res[0] == single and res[1] == single
    for pk in pks:
        pk.close()
    print("distprove ok", flush=True)
print("workload done")
