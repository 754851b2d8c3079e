"""Child process of tests/test_plonk_dist_gpu.py: `world` ranks of the multi-GPU SPMD prover emulated on the GPUs this
process sees (rank k on device k % device_count; on a 1-GPU box all ranks share the device).  Every rank has its own
context, SRS and proving key and proves from its own host thread; the proofs must all equal the single-GPU proof of
rank 0's key.  Prints one JSON line."""
import json
import os
import sys
import threading
import time

# ranks that share ONE device (this emulation) spin-wait on each other inside one CUDA context, so nothing may
# synchronise that context behind a waiting rank's back: every stream gets its own hardware queue, and kernels are
# loaded eagerly (lazy loading of a kernel's first launch synchronises the context — CUDA programming guide, "Lazy
# Loading: concurrent execution").  Ranks on different GPUs / in different processes do not need either.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ["CUDA_MODULE_LOADING"] = "EAGER"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import noir_backend_using_gnark_b200 as zk  # noqa: E402
from noir_backend_using_gnark_b200 import plonk as zkp  # noqa: E402
from oracle import bn254 as o  # noqa: E402
from oracle import plonk as pl  # noqa: E402


def main() -> None:
    world, gates, nb_public = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    with_oracle = len(sys.argv) > 4 and sys.argv[4] == "oracle"
    bad_row = int(sys.argv[5]) if len(sys.argv) > 5 else -1
    ndev = zk.load().b200zk_device_count()
    alpha = o.random_fr(1, 0xB2000005)[0]
    cs_o, x = pl.synthetic_chain_circuit(gates, 0xB2000004 + gates, nb_public)
    g = cs_o.gates
    cs_p = zkp.SparseR1CS(cs_o.nb_public, cs_o.nb_secret, [t.ql for t in g], [t.qr for t in g], [t.qm for t in g],
                          [t.qo for t in g], [t.qk for t in g], [t.a for t in g], [t.b for t in g], [t.c for t in g])
    size = 1
    while size < gates + nb_public:
        size <<= 1
    ctxs = [zk.Context(k % ndev) for k in range(world)]
    srss = [zk.SRS.NewSRS(size + 3, o.fr_to_mont_bytes([alpha]), c).precompute() for c in ctxs]
    pks = [zkp.ProvingKey.Setup(cs_p, s, c) for s, c in zip(srss, ctxs)]
    st = pl.BlindingStream(0xB2000006)
    blind = np.frombuffer(b"".join(o.limbs_le(st.next_mont()) for _ in range(9)), dtype=np.uint8)
    sol = np.frombuffer(o.fr_to_mont_bytes(x), dtype=np.uint8).copy()
    if bad_row >= 0:   # corrupt one secret wire: every rank must report the same unsatisfied constraint
        sol[32 * (nb_public + bad_row)] ^= 1
    single = None
    if bad_row < 0:
        single = pks[0].Prove(sol, blind).blob
    zkp.ProvingKey.JoinLocal(pks)
    out = {"world": world, "gates": gates, "devices": ndev}
    for rep in range(2):   # twice: the second proof reuses exchange buffers, barrier epochs and partial slots
        res, errs = [None] * world, [None] * world

        def run(k):
            try:
                res[k] = pks[k].Prove(sol if k == 0 else None, blind if k == 0 else None).blob
            except BaseException as e:  # noqa: BLE001
                errs[k] = repr(e)

        th = [threading.Thread(target=run, args=(k,)) for k in range(world)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        out["ms_%d" % rep] = (time.perf_counter() - t0) * 1e3
        out["errors_%d" % rep] = errs
        if bad_row < 0:
            out["all_equal_single_%d" % rep] = all(r == single for r in res)
    if with_oracle and bad_row < 0:
        srs_o = pl.SRS(size + 3, alpha)
        pk_o = pl.setup(cs_o, srs_o)
        want = pl.prove(cs_o, pk_o, srs_o, x, pl.BlindingStream(0xB2000006)).to_bytes()
        out["equals_oracle"] = zkp.Proof(res[world - 1]).to_gnark_bytes() == want
        out["verified"] = bool(pl.verify(pl.Proof.from_bytes(want), pk_o.vk, x[:nb_public], srs_o.g2))
    for pk in pks:
        pk.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
