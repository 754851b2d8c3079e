#!/usr/bin/env python
"""Benchmark of the PLONK hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload msm|ntt]

Headline workload (BASELINE.json: "G1 MSM Mpoints/s @2^24"): one step = one BN254 G1 MSM of 2^24 points per GPU
(KZG commitment of a 2^24-coefficient polynomial against a device-resident SRS shard).  Under torchrun every rank
owns the contiguous point range [rank*2^24, (rank+1)*2^24) of one SRS; a step ends with an NCCL all-gather of the
ranks' 128-byte partial sums and the final addition, so the N-GPU job computes ONE MSM of N*2^24 points (weak
scaling, no data-path collective besides that 128-byte exchange).
`value`   : points/s over all ranks with scalars and bases resident in HBM.
`e2e`     : same, through the C-ABI host entry point b200zk_msm_g1 (host scalars, H2D + D2H inside the timed region).
`roofline`: msm_accumulate_kernel (bucket accumulation, >90 % of the step) against the integer multiply-add
            roofline: algorithmic 21 760 32x32-bit MACs per point / CUDA-event time of that kernel,
            peak = IMAD.WIDE rate measured live by b200zk_microbench.  `ntt` carries the fr NTT HBM figures.
`cpu_baseline` / `--impl reference`: the C restatement of gnark's MultiExp (oracle/bn254_ref.c, pthreads on all host
            cores) on a bounded sample — gnark itself (Go) cannot run in this image; see DESIGN.md.
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
CPU_SAMPLE_LOG2 = 20


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def random_fr_images(n: int, seed: int):
    """n uniform-ish fr elements as in-memory (Montgomery) images: 4 x u64 limbs, top limb < 2^60 (< r).
    Input synthesis for the B200 arm — deliberately NOT the oracle's generator."""
    import numpy as np

    rng = np.random.default_rng(seed)
    limbs = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    limbs[:, 3] &= (1 << 60) - 1
    return limbs.view(np.uint8).reshape(-1)


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
# reference arm / cpu baseline: the C restatement of gnark's CPU MultiExp on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_msm_sample(points, scalars, n: int, reps: int):
    from oracle import cref

    cores = cref.ncores()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cref.msm(points, scalars, n, nthreads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, cores


def cpu_bases(n: int):
    """Structured bases P_i = (a + i*b)G built on the CPU (the reference arm must not touch the GPU library)."""
    from oracle import bn254 as o
    from oracle import cref

    a, b = SEED_SRS % 9973, 7919
    return cref.g1_arith_progression(o.g1_to_bytes([o.g1_mul(o.G1_GEN, a)]), o.g1_to_bytes([o.g1_mul(o.G1_GEN, b)]), n)


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    from oracle import cref

    n = 1 << CPU_SAMPLE_LOG2
    pts = cpu_bases(n)
    sc = cref.random_fr(n, SEED_SCALARS)
    cores = cref.ncores()
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cref.msm(pts, sc, n, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.msm(pts, sc, n, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
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
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery)",
        "data": "synthetic",
        "config": {"workload": "G1 MSM 2^24 points/GPU (KZG commit), bounded sample 2^%d points per step" % CPU_SAMPLE_LOG2,
                   "log2n": 24, "sample_log2n": CPU_SAMPLE_LOG2},
        "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": "2^%d-point MultiExp, C restatement of gnark-crypto multiexp (not gnark itself: no Go toolchain)" % CPU_SAMPLE_LOG2},
        "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def prove_cpu_baseline(ctx, sample_log2: int) -> dict:
    """CPU arm of the prove metric on a bounded sample: the C port of the prover (oracle/bn254_ref.c, all host
    cores) on a 2^sample_log2-row chain circuit, next to the CUDA prover on the SAME circuit, SRS and blinding —
    whose proof bytes must be identical.  The C port is handed the key polynomials of the device key (what gnark holds
    after ProvingKey.ReadFrom) and derives the coset forms itself; setup parity is covered by the tests."""
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
    return {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port",
            "sample": "full PLONK prove of a 2^%d-row chain circuit by the C port of the prover (not gnark: no Go "
                      "toolchain here)" % sample_log2,
            "b200_ms_same_circuit": gpu_ms, "proof_bytes_identical_to_cpu_port": bool(same)}


def run_prove(ctx, log2n: int, rank: int = 0, world: int = 1, cpu_sample_log2: int = 0):
    """BASELINE metric 1: full PLONK prove latency (device-resident prover, b200zk_plonk_prove) on the synthetic
    chain circuit of 2^log2n - 1 gates + 1 public input; the proof is checked by the independent verifier.
    With world > 1 the prover's commitments are sharded by point range over the ranks (dist_prove.py): rank 0
    proves, the other ranks serve MSM shards; both the single-GPU and the sharded latency are reported."""
    import numpy as np

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200 import plonk as zkp

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    n = 1 << log2n
    a_int = SEED_SRS * 0x9E3779B97F4A7C15 % zkp.R_MOD
    a_img = zkp.fr_to_mont([a_int])
    com = None
    if world > 1:
        from noir_backend_using_gnark_b200.dist_prove import ShardedCommitter

        m = -(-(n + 3) // world)
        shard = zk.SRS.NewSRS(m, a_img, ctx, first=rank * m).precompute()
        com = ShardedCommitter(ctx, shard, m)
        if rank != 0:
            com.serve()
            shard.close()
            return None
    from prove_bench import synthetic

    c = synthetic(log2n)
    srs = zk.SRS.NewSRS(n + 3, a_img, ctx).precompute()
    t0 = time.perf_counter()
    pk = zkp.ProvingKey.SetupRaw(srs, log2n, log2n + 2, 1, c["nb_wires"], c["ql"], c["qr"], c["qm"], c["qo"], c["qk"],
                                 c["lro"], ctx)
    setup_ms = (time.perf_counter() - t0) * 1e3
    blind = random_fr_images(9, 0xB2000006)
    import torch

    sol_pinned = torch.from_numpy(c["sol"]).pin_memory()   # the caller's host buffer, page-locked (b200zk_host_alloc for C hosts)
    c["sol"] = sol_pinned.numpy()

    def timed():
        pk.Prove(c["sol"], blind)  # warm-up
        ts = []
        l0 = ctx.launch_count
        for _ in range(3):
            t0 = time.perf_counter()
            pr = pk.Prove(c["sol"], blind)
            ts.append((time.perf_counter() - t0) * 1e3)
        return pr, ts, (ctx.launch_count - l0) // 3

    sharded = None
    try:
        proof, times, launches = timed()
        if com is not None:
            com.attach(pk)
            proof_s, times_s, _ = timed()
            com.detach(pk)
            sharded = {"gpus": world, "ms": min(times_s), "ms_all": times_s,
                       "same_proof_bytes": proof_s.blob == proof.blob,
                       "error": repr(com.error) if com.error else None}
    finally:
        if com is not None:
            com.stop()   # always release the worker ranks, also when rank 0 failed
    # acceptance by the independent verifier (checker only: oracle/plonk.py pairing check)
    from oracle import bn254 as o
    from oracle import plonk as pl

    S = [o.g1_from_bytes(b)[0] for b in pk.vk_points]
    vk = pl.VerifyingKey(n, pow(n, -1, zkp.R_MOD), o.Domain(n).generator, 1, 5, S[:3], S[3], S[4], S[5], S[6], S[7])
    ok = bool(pl.verify(pl.Proof.from_bytes(proof.to_gnark_bytes()), vk, [c["x0"]], (pl.G2_GEN, pl.g2_mul(pl.G2_GEN, a_int))))
    pk.close()
    srs.close()
    cpu = prove_cpu_baseline(ctx, cpu_sample_log2) if cpu_sample_log2 else None
    return {"metric": "plonk_prove_latency", "log2_gates": log2n, "ms": min(times), "ms_all": times, "unit": "ms",
            "cpu_baseline": cpu,
            "higher_is_better": False, "setup_ms": setup_ms, "launches_per_prove": int(launches), "verified": ok,
            "api": "b200zk_plonk_prove (host solution vector in, 832-byte proof out; H2D/D2H included)",
            "h2d_bytes": n * 32 + 288, "d2h_bytes": 832, "sharded_msm": sharded}


def run_b200(args, rank: int, world: int, local_rank: int) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import noir_backend_using_gnark_b200 as zk
    from noir_backend_using_gnark_b200 import plonk as zkp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = zk.Context(local_rank)
    ext = ctx.torch_stream()
    peaks = load_peaks()

    log2n = args.log2n
    n = 1 << log2n
    alpha_img = zkp.fr_to_mont([SEED_SRS * 0x9E3779B97F4A7C15 % zkp.R_MOD])
    srs = zk.SRS.NewSRS(n, alpha_img, ctx, first=rank * n)
    precompute_s = None
    if not args.no_precompute:
        # one-time, like the SRS upload itself: the bases are static across commitments
        t0 = time.perf_counter()
        srs.precompute()
        precompute_s = time.perf_counter() - t0
    srs_windows = srs.windows(n)
    h_sc = torch.from_numpy(random_fr_images(n, SEED_SCALARS + rank)).pin_memory()
    d_sc = h_sc.to(dev)
    torch.cuda.synchronize()

    part = torch.zeros(128, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(128 * world, dtype=torch.uint8, device=dev)
    result = torch.zeros(64, dtype=torch.uint8, device=dev)

    def step():
        with torch.cuda.stream(ext):
            if world == 1:
                zk.MultiExp(srs, d_sc, n=n, out=result)
            else:
                zk.MultiExp(srs, d_sc, n=n, out=part, partial=True)
                dist.all_gather_into_tensor(gathered, part)
                zk.SumPartials(ctx, gathered, out=result)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

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
    value = world * n / (ms_step * 1e-3) / 1e6

    # ---- correctness of what was timed: closed form Commit(p) = p(alpha)*G is checked in tests at 2^22; here the
    # single-GPU result must equal the host-API result below (same inputs, different entry point).
    res_dev = result.cpu().numpy().tobytes()

    # ---- e2e: host scalars through the C ABI (H2D + MSM + D2H per step)
    e2e_steps = max(1, min(args.steps, 5))
    if world == 1:
        zk.MultiExp(srs, h_sc, n=n)   # untimed: the first host-scalar call allocates the 512 MiB staging buffer
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if world == 1:
            res_host = zk.MultiExp(srs, h_sc, n=n)
        else:
            with torch.cuda.stream(ext):
                d_tmp = h_sc.to(dev, non_blocking=True)
                zk.MultiExp(srs, d_tmp, n=n, out=part, partial=True)
                dist.all_gather_into_tensor(gathered, part)
                zk.SumPartials(ctx, gathered, out=result)
                res_host = result.cpu().numpy().tobytes()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert res_host == res_dev, "host-API and device-API MSM results differ"

    # ---- PLONK prove (all ranks take part when world > 1: the commitments are sharded)
    prove_info = None
    cpu_pts = None
    if rank == 0 and not args.no_cpu:
        cpu_pts = np.frombuffer(srs.download(0, 1 << min(CPU_SAMPLE_LOG2, log2n)), dtype=np.uint8)
    if not args.no_prove:
        srs.close()          # free the 2^24 window table before the prover allocates its arena
        del d_sc
        torch.cuda.empty_cache()
        prove_info = run_prove(ctx, args.prove_log2n, rank, world, 0 if args.no_cpu else 20)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

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
    achieved = MACS_PER_POINT * n / (acc_ms_per * 1e-3) if acc_ms_per > 0 else 0.0
    # MACs the kernel really issues: one mixed XYZZ addition = 8 multiplications (136 MACs) + 2 squarings (100 MACs);
    # additions per point = number of windows (12 with the 2^24 window table, 16 classic windows)
    adds_per_point = srs_windows
    executed_macs = (adds_per_point * (8 * 136 + 2 * 100) * n / (acc_ms_per * 1e-3)) if (adds_per_point and acc_ms_per > 0) else None
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
        if not args.no_cpu:
            # CPU arm of the NTT metric on a bounded sample: the C port of fft.Domain.FFT on all host cores
            from oracle import cref

            slog = min(nlog, 22)
            sample = random_fr_images(1 << slog, SEED_NTT)
            cores = cref.ncores()
            cref.ntt(sample, slog, 0, 0, 0, cores)          # builds the domain tables
            t0 = time.perf_counter()
            cref.ntt(sample, slog, 0, 0, 0, cores)
            dt = time.perf_counter() - t0
            ntt_info["cpu_baseline"] = {"value": 64.0 * (1 << slog) / dt / 1e9, "unit": "GB/s", "ms": dt * 1e3, "cores": cores,
                                        "kind": "port", "sample": "one 2^%d DIF transform (includes copying the input), C "
                                        "restatement of gnark-crypto fft.Domain.FFT (not gnark itself)" % slog}

    cpu = None
    if not args.no_cpu:
        ns = 1 << min(CPU_SAMPLE_LOG2, log2n)
        pts = cpu_pts  # first 2^20 SRS points, downloaded above; this leg is the only oracle compute in this arm
        dt, cores = cpu_msm_sample(pts, h_sc.numpy()[: ns * 32], ns, reps=1)
        cpu = {"value": ns / dt / 1e6, "unit": "Mpoints/s", "cores": cores, "kind": "port",
               "sample": "first 2^%d points of the same MSM, C restatement of gnark-crypto MultiExp on all host cores "
                         "(gnark itself needs Go: not runnable here)" % min(CPU_SAMPLE_LOG2, log2n)}

    line = {
        "metric": "bn254_g1_msm_throughput",
        "value": value,
        "unit": "Mpoints/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32x8 (254-bit Montgomery fp/fr)",
        "data": "synthetic",
        "config": {"workload": "G1 MSM 2^%d points per GPU vs device-resident KZG SRS shard (kzg.Commit)" % log2n,
                   "log2n": log2n, "points_total": world * n, "sharding": "point range per rank, 128 B all-gather",
                   "l2": "inputs (1.5 GiB/GPU) larger than L2, no flush needed",
                   "msm_mode": "classic windows" if args.no_precompute else
                   "precomputed window multiples of the static SRS (one-time %.2f s, not in the timed region)" % precompute_s,
                   "seed_scalars": hex(SEED_SCALARS),
                   "seed_srs": hex(SEED_SRS)},
        "clocks": clocks,
        "e2e": {"value": world * n / e2e_s / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": n * 32,
                "d2h_bytes_per_step": 64, "ms_per_step": e2e_s * 1e3,
                "api": "b200zk_msm_g1 (host scalars)" if world == 1 else "pinned H2D + b200zk_msm_g1_dev + all-gather + D2H"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "int-mad (fma-pipe IMAD.WIDE; MSM is not HBM- or tensor-bound)", "kernel": "msm_accumulate_kernel",
                     "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TMAC/s",
                     "frac": achieved / imad_peak if imad_peak else None,
                     "note": "achieved = SURVEY 8d algorithmic unit (16 windows x 10 modmul x 136 MACs = 21760 MACs/point) / kernel "
                             "time; with the window table the kernel executes 12 additions per point (8 mul x 136 + 2 sqr x 100 "
                             "MACs each), so the algorithmic figure can exceed the pipe peak; frac_executed is the executed-MAC rate",
                     "achieved_executed": (executed_macs / 1e12) if executed_macs else None,
                     "frac_executed": (executed_macs / imad_peak) if (executed_macs and imad_peak) else None,
                     "traffic": (load_traffic().get("msm_accumulate_kernel@2^24_table", {}).get("bytes")
                                 if (log2n == 24 and not args.no_precompute) else None),
                     "peak_source": "b200zk_microbench IMAD.WIDE.U32, measured in this run",
                     "kernel_ms": acc_ms_per, "kernel_share_of_step": phase_share.get("msm_accumulate"),
                     "fp_mul_per_s_peak": fpmul_peak,
                     "hbm_bytes_algorithmic": 96 * n},
        "phase_share": phase_share,
        "ntt": ntt_info,
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
