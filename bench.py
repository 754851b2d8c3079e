#!/usr/bin/env python
"""Benchmark of the PLONK hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU)

Headline workload (BASELINE.json: "G1 MSM Mpoints/s @2^24"): one step = ONE BN254 G1 MSM of 2^24 points (KZG
commitment of a 2^24-coefficient polynomial against a device-resident SRS).  With N ranks the SAME 2^24-point MSM is
sharded by point range: rank k owns SRS points [k*2^24/N, (k+1)*2^24/N) (and their window table) and the matching
scalars; a step ends with an NCCL all-gather of the 128-byte partial sums and the final addition — strong scaling
(BASELINE config 4).  The result is asserted equal to the one-GPU MSM of the same inputs in the same run.
`value`        : points/s with scalars and bases resident in HBM (CUDA events, max over ranks).
`e2e`          : same MSM through the C-ABI host entry points (b200zk_msm_g1 / b200zk_msm_g1_shard): pinned HOST scalars,
                 H2D + D2H inside the timed region.
`roofline`     : msm_accumulate_kernel (bucket accumulation) against the integer multiply-add roofline; `frac` = executed
                 32x32 MACs / measured IMAD.WIDE peak, `frac_algorithmic` = SURVEY §8d's 21 760 MACs per point / peak.
`msm_classic`  : the same MSM without the precomputed window table (memory-neutral figure).
`weak`         : (N > 1) 2^24 points PER GPU, the round-1 headline, kept for continuity.
`ntt`          : fr NTT 2^24 (HBM figures);  `dist_ntt` (N > 1): four-step NTT at 2^24 / 2^26 (/ 2^28 at N = 8), NCCL
                 all-to-all and fused peer-store exchange, asserted bit-exact against the one-GPU transform.
`plonk_prove`  : full prove latency at 2^22 gates; N > 1: the SPMD multi-GPU prover (b200zk_plonk_join), same proof bytes.
`cpu_baseline` / `--impl reference`: the C restatement of gnark's MultiExp (oracle/bn254_ref.c, pthreads on all host
                 cores) on the same 2^24-point workload — gnark itself (Go) cannot run in this image; see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED_SCALARS = 0xB2000001
SEED_SRS = 0xB2000005
SEED_NTT = 0xB2000003
MACS_PER_POINT = 21760          # 16 windows x 10 modmul x 136 MACs
CPU_BUDGET_S = 200.0            # the CPU arms choose their sample so that the whole run stays within a few minutes


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def common_config(log2n: int, world: int) -> dict:
    """the workload both arms (b200 / reference) are measured on — identical dicts, so the driver can match them"""
    return {"workload": "one G1 MSM of 2^%d points (kzg.Commit against a resident SRS), sharded by point range over "
                        "the GPUs" % log2n,
            "log2n": log2n, "points_total": 1 << log2n,
            "scalars": "uniform 254-bit, seed %s" % hex(SEED_SCALARS),
            "bases": "KZG SRS powers alpha^i*G (b200 arm) / arithmetic progression (a+i*b)*G (CPU arm), seed %s" % hex(SEED_SRS),
            "l2": "inputs (>= 1.5 GiB) larger than L2, no flush needed"}


def random_fr_images(n: int, seed: int):
    """n uniform-ish fr elements as in-memory (Montgomery) images: 4 x u64 limbs, top limb < 2^60 (< r).
    Input synthesis for the B200 arm — deliberately NOT the oracle's generator."""
    import numpy as np

    rng = np.random.default_rng(seed)
    limbs = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    limbs[:, 3] &= (1 << 60) - 1
    return limbs.view(np.uint8).reshape(-1)


def random_fr_device(n: int, seed: int, dev):
    """the same kind of input generated on the device (identical on every rank for one seed): for the 2^26..2^28 NTTs"""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    limbs = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device=dev, generator=g)
    limbs[:, 3] &= (1 << 60) - 1
    return limbs.view(torch.uint8).reshape(-1)


def load_traffic() -> dict:
    """per-launch DRAM bytes of the dominant kernels from the committed ncu captures (profiles/ncu_traffic.json)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def load_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def pin_near_gpu(index: int):
    """run this process (and first-touch its pinned buffers) on the CPUs of the GPU's NUMA node; returns the previous
    affinity so that the CPU-baseline legs can take all cores back"""
    try:
        import pynvml

        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (NVML, 20 ms period)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is None:
            return
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the C restatement of gnark's CPU MultiExp
# ------------------------------------------------------------------------------------------------------
def cpu_bases(n: int):
    """Structured bases P_i = (a + i*b)G built on the CPU (the reference arm must not touch the GPU library); the
    progression is cut into one piece per core (each piece starts from its own scalar multiple)."""
    import numpy as np

    from oracle import bn254 as o
    from oracle import cref

    a, b = SEED_SRS % 9973, 7919
    cores = max(1, min(cref.ncores(), 32))
    per = -(-n // cores)
    step = o.g1_to_bytes([o.g1_mul(o.G1_GEN, b)])
    out = [None] * cores

    def piece(t):
        lo, hi = t * per, min(n, (t + 1) * per)
        if hi > lo:
            out[t] = cref.g1_arith_progression(o.g1_to_bytes([o.g1_mul(o.G1_GEN, a + lo * b)]), step, hi - lo)

    th = [threading.Thread(target=piece, args=(t,)) for t in range(cores)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return np.concatenate([x for x in out if x is not None])


def cpu_sample_log2(log2n: int, steps: int) -> tuple:
    """largest sample (<= the workload) whose `steps` repetitions fit the CPU budget, from a 2^18-point probe"""
    from oracle import cref

    cores = cref.ncores()
    pl = 18
    pts = cpu_bases(1 << pl)
    sc = cref.random_fr(1 << pl, SEED_SCALARS)
    t0 = time.perf_counter()
    cref.msm(pts, sc, 1 << pl, nthreads=cores)
    probe = time.perf_counter() - t0
    lg = log2n
    # Pippenger at c = 16: time ~ proportional to n above 2^20; the probe at 2^18 overestimates the per-point cost
    while lg > 20 and probe * (1 << (lg - pl)) * 0.7 * steps > CPU_BUDGET_S:
        lg -= 1
    return lg, cores


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    from oracle import cref

    log2n = args.log2n
    lg, cores = cpu_sample_log2(log2n, args.steps + args.warmup)
    n = 1 << lg
    pts = cpu_bases(n)
    sc = cref.random_fr(n, SEED_SCALARS)
    for _ in range(args.warmup):
        cref.msm(pts, sc, n, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.msm(pts, sc, n, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    sample = ("the whole 2^%d-point MultiExp per step" % lg if lg == log2n else
              "2^%d of the 2^%d points per step (CPU budget %.0f s for %d steps)" % (lg, log2n, CPU_BUDGET_S, args.steps + args.warmup))
    line = {
        "impl": "reference",
        "metric": "bn254_g1_msm_throughput",
        "value": val,
        "unit": "Mpoints/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery fp/fr)",
        "data": "synthetic",
        "config": common_config(log2n, world),
        "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": sample + "; C restatement of gnark-crypto multiexp (oracle/bn254_ref.c), not gnark itself: no "
                                            "Go toolchain in this image",
                         "sample_log2n": lg},
        "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# B200 arm: PLONK prove
# ------------------------------------------------------------------------------------------------------
def prove_cpu_baseline(ctx, sample_log2: int, target_log2: int) -> dict:
    """CPU arm of the prove metric: the C port of the prover (oracle/bn254_ref.c, all host cores) on a
    2^sample_log2-row chain circuit, next to the CUDA prover on the SAME circuit, SRS and blinding — whose proof bytes
    must be identical.  The C port is handed the key polynomials of the device key (what gnark holds after
    ProvingKey.ReadFrom) and derives the coset forms itself; setup parity is covered by the tests."""
    import numpy as np

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200 import plonk as zkp
    from oracle import cref
    from oracle import plonk as pl

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from prove_bench import synthetic

    n = 1 << sample_log2
    c = synthetic(sample_log2)
    alpha = SEED_SRS * 0x9E3779B97F4A7C15 % zkp.R_MOD
    srs_d = zk.SRS.NewSRS(n + 3, zkp.fr_to_mont([alpha]), ctx).precompute()
    pk_d = zkp.ProvingKey.SetupRaw(srs_d, sample_log2, sample_log2 + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"],
                                   c["qk"], c["lro"], ctx)
    blind = random_fr_images(9, 0xB2000006)
    sol = np.ascontiguousarray(c["sol"])
    pk_d.Prove(sol, blind)
    t0 = time.perf_counter()
    proof_gpu = pk_d.Prove(sol, blind)
    gpu_ms = (time.perf_counter() - t0) * 1e3
    cores = cref.ncores()
    cp = pl.CProver.from_arrays(sample_log2, sample_log2 + 2, 1, c["nb_wires"], [pk_d.poly(i) for i in range(9)],
                                pk_d.permutation, c["lro"], b"".join(pk_d.vk_points),
                                np.frombuffer(srs_d.download(), dtype=np.uint8), cores)
    t0 = time.perf_counter()
    blob_cpu = cp.prove_blob(sol, blind.tobytes(), cores)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = proof_gpu.blob == blob_cpu
    pk_d.close()
    srs_d.close()
    scale = float(1 << (target_log2 - sample_log2)) * target_log2 / sample_log2 if target_log2 > sample_log2 else 1.0
    out = {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port",
           "sample": "full PLONK prove of a 2^%d-row chain circuit by the C port of the prover (not gnark: no Go "
                     "toolchain here)" % sample_log2,
           "sample_log2_gates": sample_log2,
           "b200_ms_same_circuit": gpu_ms, "proof_bytes_identical_to_cpu_port": bool(same)}
    if target_log2 > sample_log2:
        out["extrapolated_ms_at_2^%d" % target_log2] = cpu_ms * scale
        out["extrapolation"] = "x %.2f = (n log n) ratio between 2^%d and 2^%d rows" % (scale, target_log2, sample_log2)
    return out


def run_prove(ctx, log2n: int, rank: int = 0, world: int = 1, cpu: bool = False):
    """BASELINE metric 1: full PLONK prove latency (device-resident prover, b200zk_plonk_prove) on the synthetic
    chain circuit of 2^log2n - 1 gates + 1 public input; the proof is checked by the independent verifier.
    With world > 1 every rank sets up the same key on its GPU and the ranks join into the SPMD multi-GPU prover
    (b200zk_plonk_join): rank 0 supplies the solution, all ranks return the same proof; both the one-GPU latency (rank 0
    alone) and the N-GPU latency (max over ranks) are reported."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200 import plonk as zkp

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from prove_bench import synthetic

    n = 1 << log2n
    a_int = SEED_SRS * 0x9E3779B97F4A7C15 % zkp.R_MOD
    a_img = zkp.fr_to_mont([a_int])
    c = synthetic(log2n)
    srs = zk.SRS.NewSRS(n + 3, a_img, ctx).precompute()
    t0 = time.perf_counter()
    pk = zkp.ProvingKey.SetupRaw(srs, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"],
                                 c["lro"], ctx)
    setup_ms = (time.perf_counter() - t0) * 1e3
    blind = random_fr_images(9, 0xB2000006)
    sol_pinned = torch.from_numpy(c["sol"]).pin_memory()   # the caller's host buffer, page-locked (b200zk_host_alloc for C hosts)
    c["sol"] = sol_pinned.numpy()

    def timed(joined: bool):
        leader = rank == 0
        sol, bl = (c["sol"], blind) if (leader or not joined) else (None, None)
        pk.Prove(sol, bl)  # warm-up
        ts = []
        l0 = ctx.launch_count
        pr = None
        for _ in range(3):
            if joined:
                dist.barrier()
            t0 = time.perf_counter()
            pr = pk.Prove(sol, bl)
            ts.append((time.perf_counter() - t0) * 1e3)
        return pr, ts, (ctx.launch_count - l0) // 3

    proof = times = launches = None
    if rank == 0:
        proof, times, launches = timed(False)
    sharded = None
    if world > 1:
        dist.barrier()
        pk.Join()
        proof_s, times_s, launches_s = timed(True)
        pk.Leave()
        t = torch.tensor(times_s, dtype=torch.float64, device="cuda:%d" % ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)            # per repetition: the slowest rank
        blobs = [None] * world
        dist.all_gather_object(blobs, proof_s.blob)
        if rank == 0:
            ms_all = [float(x) for x in t.tolist()]
            sharded = {"gpus": world, "ms": min(ms_all), "ms_all": ms_all, "launches_per_prove_rank0": int(launches_s),
                       "same_proof_bytes": all(b == proof.blob for b in blobs),
                       "what": "SPMD prover (b200zk_plonk_join): MSMs by point range, 4n-domain NTTs four-step with fused "
                               "peer-store exchange, quotient by index range; n-sized stages replicated; wall clock around "
                               "b200zk_plonk_prove, max over ranks"}
            assert sharded["same_proof_bytes"], "multi-GPU proof differs from the one-GPU proof"
    if rank != 0:
        pk.close()
        srs.close()
        return None
    # acceptance by the independent verifier (checker only: oracle/plonk.py pairing check)
    from oracle import bn254 as o
    from oracle import plonk as pl

    S = [o.g1_from_bytes(b)[0] for b in pk.vk_points]
    vk = pl.VerifyingKey(n, pow(n, -1, zkp.R_MOD), o.Domain(n).generator, 1, 5, S[:3], S[3], S[4], S[5], S[6], S[7])
    ok = bool(pl.verify(pl.Proof.from_bytes(proof.to_gnark_bytes()), vk, [c["x0"]], (pl.G2_GEN, pl.g2_mul(pl.G2_GEN, a_int))))
    assert ok, "proof rejected by the independent verifier"
    cpu_info = None
    if cpu:
        # CPU arm on the same circuit size when the budget allows (the C port needs ~35-80 s at 2^22), else 2^20 scaled
        from oracle import cref

        same_size = cref.ncores() >= 24 or log2n <= 20
        if same_size:
            cp = pl.CProver.from_arrays(log2n, log2n + 2, 1, c["nb_wires"], [pk.poly(i) for i in range(9)], pk.permutation,
                                        c["lro"], b"".join(pk.vk_points), np.frombuffer(srs.download(), dtype=np.uint8),
                                        cref.ncores())
            t0 = time.perf_counter()
            blob_cpu = cp.prove_blob(np.ascontiguousarray(c["sol"]), blind.tobytes(), cref.ncores())
            cpu_ms = (time.perf_counter() - t0) * 1e3
            cpu_info = {"value": cpu_ms, "unit": "ms", "cores": cref.ncores(), "kind": "port",
                        "sample": "the same 2^%d-row circuit, full PLONK prove by the C port of the prover (not gnark: no Go "
                                  "toolchain here)" % log2n, "sample_log2_gates": log2n,
                        "proof_bytes_identical_to_cpu_port": bool(blob_cpu == proof.blob)}
            del cp
    pk.close()
    srs.close()
    if cpu and cpu_info is None:
        cpu_info = prove_cpu_baseline(ctx, 20, log2n)
    return {"metric": "plonk_prove_latency", "log2_gates": log2n, "ms": min(times), "ms_all": times, "unit": "ms",
            "cpu_baseline": cpu_info,
            "higher_is_better": False, "setup_ms": setup_ms, "launches_per_prove": int(launches), "verified": ok,
            "api": "b200zk_plonk_prove (host solution vector in, 832-byte proof out; H2D/D2H included)",
            "h2d_bytes": n * 32 + 288, "d2h_bytes": 832, "sharded": sharded}


# ------------------------------------------------------------------------------------------------------
# B200 arm: four-step NTT over the ranks (N > 1)
# ------------------------------------------------------------------------------------------------------
def run_dist_ntt(ctx, rank: int, world: int, sizes) -> dict:
    """BASELINE config 5: fr NTT >= 2^24 as a four-step transform with the transpose over NVLink, NCCL all-to-all and the
    fused peer-store exchange, DIF forward; every rank checks its shard of the result bit for bit against the one-GPU
    transform of the same input (computed locally), and the flags are all-reduced."""
    import torch
    import torch.distributed as dist

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200.dist_ntt import DistributedDomain

    dev = torch.device("cuda", ctx.device)
    ext = ctx.torch_stream()
    out = {"variant": "FFT DIF plain: column-block shards in, row-block (contiguous) shards of the bit-reversed output",
           "nvlink_peak_gbs": 770.0, "sizes": {}}
    for log2n in sizes:
        N = 1 << log2n
        full = random_fr_device(N, SEED_NTT + log2n, dev)
        res = {}
        for p2p in (False, True):
            d = DistributedDomain(N, ctx, p2p=p2p)
            lay = d.layout
            shard0 = full.view(lay.R, lay.C, 32)[:, rank * lay.C_loc:(rank + 1) * lay.C_loc].contiguous().view(-1)
            x = shard0.clone()
            torch.cuda.synchronize()       # torch's stream wrote x; the library works on its own stream
            y = d.FFT(x, zk.DIF, False)
            ctx.sync()
            got = y.clone()
            reps = 5
            for _ in range(2):
                x.copy_(shard0)
                d.FFT(x, zk.DIF, False)
            ctx.sync()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(reps):
                x.copy_(shard0)
                torch.cuda.synchronize()
                dist.barrier()
                e0.record(ext)
                d.FFT(x, zk.DIF, False)
                e1.record(ext)
                ctx.sync()
                tot += e0.elapsed_time(e1)
            t = torch.tensor([tot / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            key = "fused_peer_stores" if p2p else "nccl_all_to_all"
            res[key + "_ms"] = float(t.item())
            if not p2p:
                # the exchange alone (NCCL all_to_all_single of the same buffers)
                tmp = torch.empty_like(x)
                with torch.cuda.stream(ext):
                    dist.all_to_all_single(tmp, x)
                ctx.sync()
                dist.barrier()
                e0.record(ext)
                with torch.cuda.stream(ext):
                    for _ in range(reps):
                        dist.all_to_all_single(tmp, x)
                e1.record(ext)
                ctx.sync()
                t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sent = 32.0 * lay.local * (world - 1) / world
                res["exchange_ms"] = float(t.item())
                res["exchange_bytes_sent_per_gpu"] = sent
                res["exchange_nvlink_gbs"] = sent / (float(t.item()) * 1e-3) / 1e9
                res["exchange_frac_of_770"] = res["exchange_nvlink_gbs"] / 770.0
                del tmp
            res[key + "_got"] = got
            d.close()
            del x, y
        # the one-GPU transform of the same input, on this rank's GPU
        dom = zk.Domain(N, ctx)
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record(ext)
        dom.FFT(full, zk.DIF, False)
        t1e.record(ext)
        ctx.sync()
        want = full[rank * lay.local * 32:(rank + 1) * lay.local * 32]
        ok = torch.tensor([int(torch.equal(res.pop("nccl_all_to_all_got"), want)),
                           int(torch.equal(res.pop("fused_peer_stores_got"), want))], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        res["bit_exact_vs_single_gpu"] = bool(ok.min().item() == 1)
        res["single_gpu_ms_first_call"] = t0e.elapsed_time(t1e)
        res["algorithmic_gbs_all_gpus_fused"] = 64.0 * N / (res["fused_peer_stores_ms"] * 1e-3) / 1e9
        assert res["bit_exact_vs_single_gpu"], "sharded NTT differs from the one-GPU transform at 2^%d" % log2n
        out["sizes"]["2^%d" % log2n] = res
        del full, want
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200 import plonk as zkp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prev_affinity = pin_near_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = zk.Context(local_rank)
    ext = ctx.torch_stream()
    peaks = load_peaks()

    log2n = args.log2n
    n_total = 1 << log2n
    n = n_total // world                      # this rank's point range [rank*n, (rank+1)*n)
    alpha_img = zkp.fr_to_mont([SEED_SRS * 0x9E3779B97F4A7C15 % zkp.R_MOD])
    srs = zk.SRS.NewSRS(n, alpha_img, ctx, first=rank * n)
    precompute_s = None
    if not args.no_precompute:
        # one-time, like the SRS upload itself: the bases are static across commitments
        t0 = time.perf_counter()
        srs.precompute()
        precompute_s = time.perf_counter() - t0
    srs_windows = srs.windows(n)
    all_sc = random_fr_images(n_total, SEED_SCALARS)      # the whole scalar vector (every rank: same seed); own slice below
    h_sc = torch.from_numpy(all_sc[rank * n * 32:(rank + 1) * n * 32].copy()).pin_memory()
    d_sc = h_sc.to(dev)
    torch.cuda.synchronize()

    part = torch.zeros(128, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(128 * world, dtype=torch.uint8, device=dev)
    result = torch.zeros(64, dtype=torch.uint8, device=dev)

    def make_step(srs_, d_sc_, n_):
        def step():
            with torch.cuda.stream(ext):
                if world == 1:
                    zk.MultiExp(srs_, d_sc_, n=n_, out=result)
                else:
                    zk.MultiExp(srs_, d_sc_, n=n_, out=part, partial=True)
                    dist.all_gather_into_tensor(gathered, part)
                    zk.SumPartials(ctx, gathered, out=result)
        return step

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(ext)
        for _ in range(steps):
            step()
        e1.record(ext)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    step = make_step(srs, d_sc, n)
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local_rank)
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0 + (args.steps if world > 1 else 0)  # + NCCL all-gather kernels
    phases = ctx.profile_read()
    ctx.profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n_total / (ms_step * 1e-3) / 1e6
    res_dev = result.cpu().numpy().tobytes()

    # ---- e2e: host scalars through the C ABI (H2D + MSM + D2H per step)
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        if world == 1:
            return zk.MultiExp(srs, h_sc, n=n)
        with torch.cuda.stream(ext):
            zk.MultiExpShard(srs, h_sc, n=n, out=part)
            dist.all_gather_into_tensor(gathered, part)
            zk.SumPartials(ctx, gathered, out=result)
            return result.cpu().numpy().tobytes()

    e2e_step()   # untimed: the first host-scalar call allocates the staging buffer
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res_host = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert res_host == res_dev, "host-API and device-API MSM results differ"

    # ---- correctness of what was timed, in this run: the same MSM on ONE GPU without the window table (classic
    # c = 16 windows + Horner: a different bucket layout and summation order) must give the same 64 bytes; its timing is
    # the memory-neutral figure.  (Closed form p(alpha)*G: tests/test_msm_gpu.py, up to 2^22.)
    classic = None
    if rank == 0 and not args.no_check:
        srs_full = srs if (world == 1 and args.no_precompute) else zk.SRS.NewSRS(n_total, alpha_img, ctx)
        d_all = torch.from_numpy(all_sc).to(dev)
        one = torch.zeros(64, dtype=torch.uint8, device=dev)
        ctx.lib.b200zk_msm_set_window(ctx.handle, 16)     # forces the classic path also when a table exists

        def classic_step():
            with torch.cuda.stream(ext):
                zk.MultiExp(srs_full, d_all, n=n_total, out=one)

        for _ in range(2):
            classic_step()
        ctx.sync()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(ext)
        for _ in range(3):
            classic_step()
        c1.record(ext)
        ctx.sync()
        ctx.lib.b200zk_msm_set_window(ctx.handle, 0)
        cms = c0.elapsed_time(c1) / 3
        same = one.cpu().numpy().tobytes() == res_dev
        classic = {"ms_per_step": cms, "value": n_total / (cms * 1e-3) / 1e6, "unit": "Mpoints/s", "gpus": 1,
                   "mode": "classic 16-bit windows + Horner, no precomputed table (no extra memory)",
                   "equals_headline_result": bool(same)}
        assert same, "sharded / table-mode MSM differs from the one-GPU classic MSM"
        if srs_full is not srs:
            srs_full.close()
        del d_all
    del all_sc

    # ---- N > 1: the round-1 headline (2^24 points PER GPU, weak scaling), kept as an extra key
    weak = None
    if world > 1 and not args.no_weak:
        srs.close()
        del d_sc
        torch.cuda.empty_cache()
        srs_w = zk.SRS.NewSRS(n_total, alpha_img, ctx, first=rank * n_total)
        if not args.no_precompute:
            srs_w.precompute()
        d_w = torch.from_numpy(random_fr_images(n_total, SEED_SCALARS + 1 + rank)).to(dev)
        wms = timed(make_step(srs_w, d_w, n_total), 5, 3) / 5
        weak = {"points_per_gpu": n_total, "points_total": world * n_total, "ms_per_step": wms,
                "value": world * n_total / (wms * 1e-3) / 1e6, "unit": "Mpoints/s", "scaling": "weak"}
        srs_w.close()
        del d_w
        srs = None
        torch.cuda.empty_cache()

    # ---- N > 1: four-step NTT over the ranks
    dist_ntt = None
    if world > 1 and not args.no_ntt:
        if srs is not None:
            srs.close()
            srs = None
            torch.cuda.empty_cache()
        sizes = [24, 26] + ([28] if world == 8 else [])
        dist_ntt = run_dist_ntt(ctx, rank, world, sizes)

    # ---- PLONK prove (all ranks take part when world > 1)
    prove_info = None
    if not args.no_prove:
        if srs is not None:
            srs.close()          # free the window table before the prover allocates its arena
            srs = None
        torch.cuda.empty_cache()
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)     # the CPU baseline of the prove metric uses all host cores
        prove_info = run_prove(ctx, args.prove_log2n, rank, world, cpu=(world == 1 and not args.no_cpu))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)

    # ---- rank 0 only: the same prover through the reference's own string FFI (include/gnark_backend_ffi.h), in a
    # child process because the FFI library keeps its own SRS / key state and reads $XDG_CONFIG_HOME
    ffi_info = None
    if not args.no_prove and world == 1:
        import subprocess

        try:
            ffi_info = {"api": "PlonkPreprocess / PlonkProveWithPK / PlonkVerifyWithVK (GoString payloads: ACIR JSON, hex felts, "
                               "hex keys); wall clock per call; 2^3 rows = the size of the reference's own test circuits "
                               "(BASELINE.json config 1)"}
            for lg in (3, 16):
                r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ffi_bench.py"), str(lg), "5"],
                                   capture_output=True, text=True, timeout=300)
                ffi_info["rows_2^%d" % lg] = (json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0
                                              else {"error": r.stderr[-300:]})
        except Exception as e:  # the headline numbers do not depend on this leg
            ffi_info = {"error": repr(e)}

    # ---- rank 0 only: roofline denominators, NTT figures, CPU baseline
    imad_peak = ctx.microbench(0)
    fpmul_peak = ctx.microbench(1)
    acc_ms, acc_cnt = phases["msm_accumulate"]
    acc_ms_per = acc_ms / max(acc_cnt, 1)
    algorithmic = MACS_PER_POINT * n / (acc_ms_per * 1e-3) if acc_ms_per > 0 else 0.0
    # MACs the kernel really issues: one mixed XYZZ addition = 6 multiplications (136 MACs) + 2 squarings (100 MACs) + one
    # two-product sweep for Y3 (fe_mul2add: 200 MACs instead of 2 x 136) = 1216;
    # additions per point = number of windows (12 with the 2^24 window table, 16 classic windows)
    MACS_PER_ADD = 6 * 136 + 2 * 100 + 200
    executed = (srs_windows * MACS_PER_ADD * n / (acc_ms_per * 1e-3)) if (srs_windows and acc_ms_per > 0) else None
    phase_share = {k: round(v[0] / max(ms_total, 1e-9), 4) for k, v in phases.items() if v[1]}

    ntt_info = None
    if not args.no_ntt:
        nlog = args.ntt_log2n
        a = torch.from_numpy(random_fr_images(1 << nlog, SEED_NTT)).to(dev)
        torch.cuda.synchronize()
        d = zk.Domain(1 << nlog, ctx)
        for _ in range(3):
            d.FFT(a, zk.DIF, False)
        ctx.sync()
        reps = 10
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(ext)
        for _ in range(reps):
            d.FFT(a, zk.DIF, False)
        f1.record(ext)
        ctx.sync()
        nms = f0.elapsed_time(f1) / reps
        gbs = 64.0 * (1 << nlog) / (nms * 1e-3) / 1e9
        hbm = peaks.get("hbm_gbs", 6650.0)
        ntraffic = load_traffic().get("ntt_pass4_kernel@2^24_passes", {}).get("bytes") if nlog == 24 else None
        ntt_info = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                    "traffic": sum(ntraffic) if ntraffic else None,
                    "note": "64*N algorithmic bytes per transform; the transform is integer-multiplier-bound "
                            "(201 M Montgomery multiplications vs 68 G/s measured peak = 2.96 ms floor), see DESIGN.md",
                    "log2n": nlog, "variant": "FFT DIF plain, in place, device resident", "ms": nms,
                    "algorithmic_gbs": gbs, "hbm_peak_gbs": hbm,
                    "hbm_frac": gbs / hbm, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                    "passes": -(-nlog // 8) if nlog > 10 else 1}
        del a
        if not args.no_cpu and world == 1:
            # CPU arm of the NTT metric at the SAME size: the C port of fft.Domain.FFT on all host cores
            from oracle import cref

            sample = random_fr_images(1 << nlog, SEED_NTT)
            cores = cref.ncores()
            cref.ntt_inplace(sample, nlog, 0, 0, 0, cores)          # builds the domain tables
            t0 = time.perf_counter()
            cref.ntt_inplace(sample, nlog, 0, 0, 0, cores)
            dt = time.perf_counter() - t0
            ntt_info["cpu_baseline"] = {"value": 64.0 * (1 << nlog) / dt / 1e9, "unit": "GB/s", "ms": dt * 1e3, "cores": cores,
                                        "kind": "port", "sample": "one 2^%d DIF transform in place (the same size), C "
                                        "restatement of gnark-crypto fft.Domain.FFT (not gnark itself)" % nlog}
            del sample

    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import cref

        lg, cores = cpu_sample_log2(log2n, 2)
        ns = 1 << lg
        pts = cpu_bases(ns)
        t0 = time.perf_counter()
        cref.msm(pts, h_sc.numpy()[: ns * 32], ns, nthreads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": ns / dt / 1e6, "unit": "Mpoints/s", "cores": cores, "kind": "port", "sample_log2n": lg,
               "sample": ("the whole 2^%d-point MultiExp, once" % lg if lg == log2n else "2^%d of the 2^%d points, once" % (lg, log2n))
               + "; C restatement of gnark-crypto MultiExp on all host cores (gnark itself needs Go: not runnable here)"}

    cfg = common_config(log2n, world)
    line = {
        "metric": "bn254_g1_msm_throughput",
        "value": value,
        "unit": "Mpoints/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery fp/fr)",
        "data": "synthetic",
        "config": cfg,
        "b200": {"points_per_gpu": n, "sharding": "point range per rank, one 128-byte all-gather per MSM",
                 "msm_mode": "classic windows" if args.no_precompute else
                 "precomputed window multiples of the static SRS shard (one-time %.2f s, not in the timed region)" % precompute_s,
                 "windows": srs_windows, "result_checked_against": "one-GPU classic-window MSM of the same inputs, in this run"},
        "clocks": clocks,
        "e2e": {"value": n_total / e2e_s / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": n * 32,
                "d2h_bytes_per_step": 64, "ms_per_step": e2e_s * 1e3,
                "api": "b200zk_msm_g1 (host scalars)" if world == 1 else
                       "b200zk_msm_g1_shard (host scalars, chunked copy under compute) + all-gather + b200zk_g1_sum_dev + D2H"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "int-mad (fma-pipe IMAD.WIDE; MSM is not HBM- or tensor-bound)", "kernel": "msm_accumulate_kernel",
                     "achieved": (executed / 1e12) if executed else None, "peak": imad_peak / 1e12, "unit": "TMAC/s",
                     "frac": (executed / imad_peak) if (executed and imad_peak) else None,
                     "note": "achieved = 32x32 multiply-accumulates the kernel executes (windows x (6 mul x 136 + 2 sqr x 100 + one "
                             "two-product sweep of 200) per point) / CUDA-event time of the kernel; frac_algorithmic uses SURVEY 8d's unit (16 windows x 10 modmul "
                             "x 136 = 21760 MACs per point) and exceeds the executed figure because the window table needs fewer "
                             "additions per point",
                     "achieved_algorithmic": algorithmic / 1e12,
                     "frac_algorithmic": algorithmic / imad_peak if imad_peak else None,
                     "traffic": (load_traffic().get("msm_accumulate_kernel@2^24_table", {}).get("bytes")
                                 if (log2n == 24 and world == 1 and not args.no_precompute) else None),
                     "peak_source": "b200zk_microbench IMAD.WIDE.U32, measured in this run",
                     "kernel_ms": acc_ms_per, "kernel_share_of_step": phase_share.get("msm_accumulate"),
                     "fp_mul_per_s_peak": fpmul_peak,
                     "hbm_bytes_algorithmic": 96 * n},
        "phase_share": phase_share,
        "msm_classic": classic,
        "weak": weak,
        "ntt": ntt_info,
        "dist_ntt": dist_ntt,
        "plonk_prove": prove_info,
        "string_ffi": ffi_info,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--ntt-log2n", type=int, default=24)
    ap.add_argument("--no-ntt", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-prove", action="store_true")
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--prove-log2n", type=int, default=22)
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
