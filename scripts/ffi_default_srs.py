"""The reference's default configuration: no srs.hex yet, 1 000 000-point SRS (backend/common.go:137).  First process
generates and saves it, second process loads it; both prove and verify an 8-row and a 2^12-row circuit."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    from noir_backend_using_gnark_b200 import ffi
    from ffi_bench import circuit

    out = {}
    for lg in (3, 12):
        js, x = circuit((1 << lg) - 1)
        t = time.perf_counter()
        pk, vk = ffi.preprocess(js, 5)
        out["preprocess_2^%d_s" % lg] = round(time.perf_counter() - t, 3)
        t = time.perf_counter()
        proof = ffi.prove_with_pk(js, x, pk)
        out["prove_2^%d_s" % lg] = round(time.perf_counter() - t, 4)
        assert ffi.verify_with_vk(js, proof, x, vk)
    print(json.dumps(out))
else:
    home = tempfile.mkdtemp(prefix="b200zk_default_srs_")
    env = dict(os.environ, XDG_CONFIG_HOME=home, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "scripts"))
    env.pop("B200ZK_SRS_SIZE", None)
    for run in ("first process (generates + saves the SRS)", "second process (loads srs.hex)"):
        t = time.perf_counter()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(run, round(time.perf_counter() - t, 2), "s total;", r.stdout.strip() or r.stderr[-300:])
    print("srs.hex bytes:", os.path.getsize(os.path.join(home, "noir-lang", "srs.hex")))
