"""torchrun --nproc-per-node N scripts/dist_prove_trace.py [LOG2N]: stage-by-stage device time of the SPMD multi-GPU
prover (B200ZK_PROVE_TRACE) on rank 0, one-GPU prove first for comparison."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import noir_backend_using_gnark_b200 as zk
from noir_backend_using_gnark_b200 import plonk as zkp
from prove_bench import synthetic
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ctx = zk.Context(lr)
c = synthetic(log2n)
srs = zk.SRS.NewSRS((1 << log2n) + 3, zkp.fr_to_mont([0x1234567890ABCDEF1234567]), ctx).precompute()
pk = zkp.ProvingKey.SetupRaw(srs, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"], c["lro"], ctx)
blind = np.frombuffer(os.urandom(9 * 32), dtype=np.uint8).copy(); blind[31::32] &= 0x0F
sol = torch.from_numpy(c["sol"].copy()).pin_memory().numpy()
if rank == 0:
    pk.Prove(sol, blind)
    os.environ["B200ZK_PROVE_TRACE"] = "1"
    sys.stderr.write("--- one GPU\n")
    pk.Prove(sol, blind)
    del os.environ["B200ZK_PROVE_TRACE"]
if world > 1:
    dist.barrier()
    pk.Join()
    args = (sol, blind) if rank == 0 else (None, None)
    pk.Prove(*args)
    dist.barrier()
    if rank in (0, world - 1):
        os.environ["B200ZK_PROVE_TRACE"] = "1"
        sys.stderr.write("--- %d GPUs, rank %d\n" % (world, rank))
    t0 = time.perf_counter()
    pk.Prove(*args)
    if rank == 0:
        sys.stderr.write("wall %.2f ms\n" % ((time.perf_counter() - t0) * 1e3))
    pk.Leave()
    dist.barrier()
    dist.destroy_process_group()
