"""Multi-GPU checks, launched with torchrun (one rank per GPU):
   torchrun --nproc-per-node G --master-addr 127.0.0.1 scripts/dist_check.py [LOG2N]
 - four-step NTT over NCCL all-to-all vs the C oracle (bit-exact), all variants, + timing
 - point-range-sharded MSM (128-byte all-gather + final sum) vs the closed form p(alpha)*G."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200.dist_ntt import DistributedDomain
from oracle import bn254 as o
from oracle import cref

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = zk.Context(local)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ok = True
VARIANTS = [(i, de, c) for i in (0, 1) for de in (zk.DIF, zk.DIT) for c in (0, 1)]
if os.environ.get("DIST_CHECK_FAST"):
    VARIANTS = [(0, zk.DIF, 1), (1, zk.DIT, 1)]

# ---------------- NTT
full = cref.random_fr(1 << log2n, 0xB2000003)
d = DistributedDomain(1 << log2n, ctx)
lay = d.layout
for inv, dec, cos in VARIANTS:
    x = torch.from_numpy(lay.scatter(full, rank, column_block=(dec == zk.DIF)).copy()).to(dev)
    torch.cuda.synchronize()
    y = (d.FFTInverse if inv else d.FFT)(x, dec, bool(cos))
    ctx.sync()
    torch.cuda.synchronize()
    shards = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(shards, y)
    if rank == 0:
        got = lay.gather([s.cpu().numpy() for s in shards], column_block=(dec == zk.DIT))
        want = cref.ntt(full, log2n, inv, dec, cos, cref.ncores())
        good = got.tobytes() == want
        ok &= good
        print("ntt 2^%d g=%d inverse=%d dec=%d coset=%d: %s" % (log2n, world, inv, dec, cos, "OK" if good else "MISMATCH"), flush=True)
# ---------------- same transforms with the fused peer-store exchange (no all-to-all call)
dp = DistributedDomain(1 << log2n, ctx, p2p=True)
for inv, dec, cos in VARIANTS:
    x = torch.from_numpy(lay.scatter(full, rank, column_block=(dec == zk.DIF)).copy()).to(dev)
    torch.cuda.synchronize()
    y = (dp.FFTInverse if inv else dp.FFT)(x, dec, bool(cos))
    ctx.sync()
    torch.cuda.synchronize()
    y = y.clone()
    shards = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(shards, y)
    if rank == 0:
        got = lay.gather([s.cpu().numpy() for s in shards], column_block=(dec == zk.DIT))
        want = cref.ntt(full, log2n, inv, dec, cos, cref.ncores())
        good = got.tobytes() == want
        ok &= good
        print("ntt-p2p 2^%d g=%d inverse=%d dec=%d coset=%d: %s" % (log2n, world, inv, dec, cos, "OK" if good else "MISMATCH"), flush=True)
x = torch.from_numpy(lay.scatter(full, rank, column_block=True).copy()).to(dev)
ext = ctx.torch_stream()
for _ in range(3):
    dp.FFT(x, zk.DIF, False)
ctx.sync(); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(ext)
for _ in range(reps):
    dp.FFT(x, zk.DIF, False)
e1.record(ext)
ctx.sync(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ntt-p2p 2^%d over %d GPUs (fused peer stores): %.3f ms  (%.1f GB/s algorithmic)" %
          (log2n, world, float(t.item()), 64.0 * (1 << log2n) / float(t.item()) / 1e6), flush=True)
dp.close()

# timing of the sharded DIF transform
x = torch.from_numpy(lay.scatter(full, rank, column_block=True).copy()).to(dev)
ext = ctx.torch_stream()
for _ in range(3):
    d.FFT(x, zk.DIF, False)
ctx.sync(); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(ext)
for _ in range(reps):
    d.FFT(x, zk.DIF, False)
e1.record(ext)
ctx.sync(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print("ntt 2^%d over %d GPUs: %.3f ms  (%.1f GB/s algorithmic, all-to-all %.1f MB/GPU)" %
          (log2n, world, ms, 64.0 * (1 << log2n) / ms / 1e6, 32.0 * (1 << log2n) / world * (world - 1) / world / 1e6), flush=True)

# ---------------- MSM
n = 1 << min(log2n, 20)
alpha = o.random_fr(1, 0xB2000005)[0]
srs = zk.SRS.NewSRS(n, o.fr_to_mont_bytes([alpha]), ctx, first=rank * n)
sc = cref.random_fr(n, 0xB2000001 + rank)
d_sc = torch.from_numpy(sc).to(dev)
torch.cuda.synchronize()
part = torch.zeros(128, dtype=torch.uint8, device=dev)
gathered = torch.zeros(128 * world, dtype=torch.uint8, device=dev)
with torch.cuda.stream(ext):
    zk.MultiExp(srs, d_sc, n=n, out=part, partial=True)
    dist.all_gather_into_tensor(gathered, part)
    res = zk.SumPartials(ctx, gathered)
ctx.sync(); torch.cuda.synchronize()
# closed form: sum over all ranks of p_k(alpha) * alpha^(k*n)
coeffs = o.fr_from_mont_bytes(sc.tobytes())
acc = 0
for cf in reversed(coeffs):
    acc = (acc * alpha + cf) % o.R_MOD
acc = acc * pow(alpha, rank * n, o.R_MOD) % o.R_MOD
tot = torch.tensor([int.from_bytes(acc.to_bytes(32, "little")[8 * i:8 * i + 8], "little") - (1 << 63) for i in range(4)],
                   dtype=torch.int64, device=dev)
allv = [torch.zeros_like(tot) for _ in range(world)]
dist.all_gather(allv, tot)
if rank == 0:
    s = 0
    for v in allv:
        s += sum((int(v[i].item()) + (1 << 63)) << (64 * i) for i in range(4))
    want = o.g1_to_bytes([o.g1_mul(o.G1_GEN, s % o.R_MOD)])
    good = res.cpu().numpy().tobytes() == want
    ok &= good
    print("msm %d x 2^%d sharded: %s" % (world, n.bit_length() - 1, "OK" if good else "MISMATCH"), flush=True)
    print("ALL OK" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
