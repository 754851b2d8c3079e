"""Child process of the FFI tests: libgnark_backend_b200.so keeps per-process state (SRS cache, keys) and terminates
the process on failure like the Go library (log.Fatal), so every scenario runs in its own interpreter.
stdin: JSON list of steps; stdout: one JSON list of results (hex strings / booleans)."""
import json
import os
import sys

from noir_backend_using_gnark_b200 import ffi


def main() -> None:
    out = []
    if os.environ.get("B200ZK_BLINDING_SEED"):   # read HERE, by the test harness: the library itself ignores the environment
        ffi.seed_blinding(int(os.environ["B200ZK_BLINDING_SEED"], 0))
    for step in json.load(sys.stdin):
        op = step["op"]
        if op == "preprocess":
            pk, vk = ffi.preprocess(step["acir"], step.get("random_value", 1))
            out.append({"pk": pk.hex(), "vk": vk.hex()})
        elif op == "prove":
            pk = step["pk"]
            if pk.startswith("@"):  # "@<i>.pk": the proving key an earlier preprocess step of this list returned
                pk = out[int(pk[1:].split(".")[0])]["pk"]
            out.append(ffi.prove_with_pk_encoded(step["acir"].encode(), ffi.encode_felts([int(v) for v in step["values"]]).encode(),
                                                 pk.encode()))
        elif op == "verify":
            out.append(ffi.verify_with_vk(step["acir"], bytes.fromhex(step["proof"]), [int(v) for v in step["values"]],
                                          bytes.fromhex(step["vk"])))
        elif op == "verify_meta":
            out.append(ffi.verify_with_meta(step["acir"], bytes.fromhex(step["proof"]), [int(v) for v in step["values"]]))
        elif op == "raw_verify":  # payload strings passed through untouched (malformed-input cases)
            lib = ffi.load_ffi()
            G = ffi.GoString.of
            out.append(int(lib.PlonkVerifyWithVK(G(step["acir"].encode()), G(step["proof"].encode()), G(step["values"].encode()),
                                                 G(step["vk"].encode()))))
        elif op == "raw_prove":  # encoded payloads passed through untouched
            pk = step["pk"]
            if pk.startswith("@"):
                pk = out[int(pk[1:].split(".")[0])]["pk"]
            out.append(ffi.prove_with_pk_encoded(step["acir"].encode(), step["values"].encode(), pk.encode()))
        elif op == "raw_preprocess":
            lib = ffi.load_ffi()
            kp = lib.PlonkPreprocess(ffi.GoString.of(step["acir"].encode()), ffi.GoString.of(step["values"].encode()))
            out.append(bool(kp.proving_key))
        else:
            raise SystemExit("unknown op " + op)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
