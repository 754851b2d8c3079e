"""torchrun --nproc-per-node G scripts/dist_prove_check.py [LOG2N]: PLONK prove with the commitments sharded over G
GPUs.  Small circuit: proof bytes must equal the oracle prover's; large circuit: timing + verifier acceptance."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import torch.distributed as dist

import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from noir_backend_using_gnark_b200.dist_prove import ShardedCommitter
from oracle import bn254 as o
from oracle import plonk as pl
from prove_bench import synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = zk.Context(local)
ALPHA = 0x1234567890ABCDEF1234567
amont = zkp.fr_to_mont([ALPHA])


def shard_srs(total: int):
    m = -(-total // world)
    return zk.SRS.NewSRS(m, amont, ctx, first=rank * m).precompute(), m


def blinding(seed):
    st = pl.BlindingStream(seed)
    return np.frombuffer(b"".join(o.limbs_le(st.next_mont()) for _ in range(9)), dtype=np.uint8)


# ---- 1. byte identity with the oracle on a small circuit (2^11 rows)
gates = 2000
cs_o, x = pl.synthetic_chain_circuit(gates, 0xB2000004, 2)
n = 2048
shard, m = shard_srs(n + 3)
com = ShardedCommitter(ctx, shard, m)
if rank == 0:
    g = cs_o.gates
    cs_p = zkp.SparseR1CS(cs_o.nb_public, cs_o.nb_secret, [t.ql for t in g], [t.qr for t in g], [t.qm for t in g],
                          [t.qo for t in g], [t.qk for t in g], [t.a for t in g], [t.b for t in g], [t.c for t in g])
    full = zk.SRS.NewSRS(n + 3, amont, ctx)
    pk = zkp.ProvingKey.Setup(cs_p, full, ctx)
    com.attach(pk)
    proof = pk.Prove(o.fr_to_mont_bytes(x), blinding(0xB2000006)).to_gnark_bytes()
    com.stop()
    assert com.error is None, com.error
    srs_o = pl.SRS(n + 3, ALPHA)
    pk_o = pl.setup(cs_o, srs_o)
    want = pl.prove(cs_o, pk_o, srs_o, x, pl.BlindingStream(0xB2000006)).to_bytes()
    print("sharded prove (%d GPUs) byte-identical to the oracle: %s" % (world, proof == want), flush=True)
    pk.close(); full.close()
else:
    com.serve()
shard.close()

# ---- 2. latency at 2^LOG2N gates
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << log2n
shard, m = shard_srs(n + 3)
com = ShardedCommitter(ctx, shard, m)
if rank == 0:
    c = synthetic(log2n)
    full = zk.SRS.NewSRS(n + 3, amont, ctx).precompute()
    pk = zkp.ProvingKey.SetupRaw(full, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"], c["lro"], ctx)
    bl = blinding(7)
    pk.Prove(c["sol"], bl)
    t_single = []
    for _ in range(3):
        t0 = time.perf_counter(); p1 = pk.Prove(c["sol"], bl); t_single.append((time.perf_counter() - t0) * 1e3)
    com.attach(pk)
    pk.Prove(c["sol"], bl)
    t_shard = []
    for _ in range(3):
        t0 = time.perf_counter(); p2 = pk.Prove(c["sol"], bl); t_shard.append((time.perf_counter() - t0) * 1e3)
    com.stop()
    assert com.error is None, com.error
    S = [o.g1_from_bytes(b)[0] for b in pk.vk_points]
    vk = pl.VerifyingKey(n, pow(n, -1, zkp.R_MOD), o.Domain(n).generator, 1, 5, S[:3], S[3], S[4], S[5], S[6], S[7])
    ok = pl.verify(pl.Proof.from_bytes(p2.to_gnark_bytes()), vk, [c["x0"]], (pl.G2_GEN, pl.g2_mul(pl.G2_GEN, ALPHA)))
    print(json.dumps({"log2_gates": log2n, "gpus": world, "prove_ms_1gpu": min(t_single), "prove_ms_sharded": min(t_shard),
                      "same_proof": p1.blob == p2.blob, "verified": bool(ok)}), flush=True)
else:
    com.serve()
dist.barrier()
dist.destroy_process_group()
