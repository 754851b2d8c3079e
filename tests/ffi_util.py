"""Helpers shared by the FFI tests (CPU and GPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_child(steps, config_home, env_extra=None, timeout=600):
    """Run tests/_ffi_child.py with $XDG_CONFIG_HOME = config_home; returns (returncode, results | None, stderr)."""
    env = dict(os.environ)
    env["XDG_CONFIG_HOME"] = str(config_home)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env.pop("B200ZK_BLINDING_SEED", None)
    env.pop("B200ZK_SRS_SIZE", None)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ffi_child.py")], input=json.dumps(steps), text=True,
                       capture_output=True, env=env, timeout=timeout, cwd=ROOT)
    res = None
    if r.returncode == 0:
        res = json.loads(r.stdout.strip().splitlines()[-1])
    return r.returncode, res, r.stderr


def write_srs_file(config_home, srs):
    from oracle import ffi_formats as ff

    d = os.path.join(str(config_home), "noir-lang")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "srs.hex"), "w") as f:
        f.write(ff.srs_bytes(srs).hex())
