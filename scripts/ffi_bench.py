"""What a Noir user sees: PlonkPreprocess -> PlonkProveWithPK -> PlonkVerifyWithVK through the reference's string FFI
(include/gnark_backend_ffi.h) on a synthetic ACIR circuit of 2^k multiplication gates (x[i+2] = x[i] * x[i+1]).
Times are wall clock around each FFI call, payload construction excluded; the JSON / hex handling inside the call is
part of what the reference's Go side does too.  Usage: python scripts/ffi_bench.py [log2_gates=16] [repeats=3]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def circuit(nb_gates: int, seed: int = 7):
    one = "%064x" % 1
    m1 = "%064x" % (R - 1)
    zero = "0" * 64
    x = [(seed * 0x9E3779B97F4A7C15 + 1) % R, (seed * 0xBF58476D1CE4E5B9 + 3) % R]
    ops = []
    for i in range(nb_gates):
        x.append(x[i] * x[i + 1] % R)
        ops.append('{"Arithmetic":{"mul_terms":[["%s",%d,%d]],"linear_combinations":[["%s",%d]],"q_c":"%s"}}' % (one, i + 1, i + 2, m1, i + 3, zero))
    js = '{"current_witness_index":%d,"opcodes":[%s],"public_inputs":[1]}' % (len(x), ",".join(ops))
    return js, x


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n_gates = (1 << lg) - 1          # + 1 public row = 2^lg rows
    home = tempfile.mkdtemp(prefix="b200zk_cfg_")
    os.environ["XDG_CONFIG_HOME"] = home
    os.environ["B200ZK_SRS_SIZE"] = str((1 << lg) + 3)
    from noir_backend_using_gnark_b200 import ffi

    js, x = circuit(n_gates)
    acir = js.encode()
    out = {"log2_rows": lg, "acir_json_MB": round(len(acir) / 1e6, 1), "host_threads": os.cpu_count()}
    import json as _json
    quoted = _json.dumps(ffi.encode_felts([12345] * len(x))).encode()
    values = ffi.encode_felts(x).encode()
    t = time.perf_counter()
    pk, vk = ffi.preprocess_encoded(acir, quoted)
    out["preprocess_first_s"] = round(time.perf_counter() - t, 3)       # includes SRS generation + srs.hex write
    t = time.perf_counter()
    pk2, vk2 = ffi.preprocess_encoded(acir, quoted)
    out["preprocess_again_s"] = round(time.perf_counter() - t, 3)
    assert pk2 == pk and vk2 == vk
    del pk2
    out["pk_hex_MB"] = round(len(pk) / 1e6, 1)
    pk, vk = pk.encode(), vk.encode()
    prove, verify = [], []
    proof = ""
    for _ in range(reps):
        t = time.perf_counter()
        proof = ffi.prove_with_pk_encoded(acir, values, pk)
        prove.append(time.perf_counter() - t)
        t = time.perf_counter()
        ok = ffi.verify_with_vk_encoded(acir, proof.encode(), values, vk)
        verify.append(time.perf_counter() - t)
        assert ok == 1
    out["prove_with_pk_s"] = [round(v, 4) for v in prove]
    out["verify_with_vk_s"] = [round(v, 4) for v in verify]
    out["proof_bytes"] = len(proof) // 2
    bad = proof[:600] + ("1" if proof[600] != "1" else "2") + proof[601:]
    assert ffi.verify_with_vk_encoded(acir, bad.encode(), values, vk) == 0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
