"""The multi-GPU SPMD prover (b200zk_plonk_join: MSMs sharded by point range, 4n-domain transforms as four-step NTTs with
the exchange fused into the butterfly pass, quotient sharded by index range) must produce the SAME proof bytes as the
single-GPU prover and as the oracle.  The ranks are emulated inside one process — one context, SRS and key per rank, one
host thread per rank — on however many GPUs the box has (the driver's GPU box has one: all ranks share it, the peer
mappings are then plain device pointers; bench.py --gpus N runs the same code over CUDA IPC on N GPUs)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_child(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_dist_prove_child.py"), *map(str, args)],
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("world,gates,nb_public", [(2, 1000, 1), (4, 1000, 3), (8, 4000, 2)])
def test_sharded_prover_equals_single_gpu_and_oracle(lib_built, world, gates, nb_public):
    out = run_child(world, gates, nb_public, "oracle")
    assert out["errors_0"] == [None] * world and out["errors_1"] == [None] * world, out
    assert out["all_equal_single_0"] and out["all_equal_single_1"], out
    assert out["equals_oracle"] and out["verified"], out


@pytest.mark.parametrize("world,log_gates", [(2, 16), (8, 18)])
def test_sharded_prover_larger_circuits(lib_built, world, log_gates):
    """sizes the Python oracle cannot reach: byte identity with the single-GPU prover (itself oracle-checked up to 2^20)"""
    out = run_child(world, (1 << log_gates) - 5, 2)
    assert out["errors_0"] == [None] * world and out["errors_1"] == [None] * world, out
    assert out["all_equal_single_0"] and out["all_equal_single_1"], out


def test_sharded_prover_reports_unsatisfied_constraint_on_every_rank(lib_built):
    out = run_child(4, 3000, 1, "no", 1234)
    errs = out["errors_0"] + out["errors_1"]
    assert all(e is not None and "UnsatisfiedConstraint" in e for e in errs), out
    assert len(set(errs)) == 1, out   # the same constraint index on every rank, both times
