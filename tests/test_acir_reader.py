"""CPU tests of the product's streaming ACIR reader + wire plan (csrc/ffi/acir_reader.h) against the oracle's
restatement of the reference's Go glue (acir/*.go, backend/common.go:45-76, backend/plonk/sparse_r1cs.go:44-107),
including its quirks: only the first mul term is read, two linear terms overwrite the mul term's wires, every
non-public value is registered once per public input, unknown witnesses map to wire 0."""
import json
import os
import random
import subprocess

import pytest

from oracle import bn254 as o
from oracle import plonk as pl

from .test_plonk_oracle import FIXTURES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = tmp_path_factory.mktemp("bin") / "acir_dump"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", str(exe), os.path.join(ROOT, "tests", "acir_dump.cpp")], check=True)
    return str(exe)


def _dump(dumper, tmp_path, js, nvalues):
    p = tmp_path / "c.json"
    p.write_text(js)
    r = subprocess.run([dumper, str(p), str(nvalues)], capture_output=True, text=True)
    return r


def _expect(js, nvalues):
    vals = list(range(1000, 1000 + nvalues))           # distinct markers: recover which value lands where
    cs, pub, sec = pl.build_sparse_r1cs(pl.decode_acir(js), vals)
    head = (cs.nb_public, cs.nb_secret, [v - 1000 for v in pub + sec])
    gates = [("%064x" % g.ql, "%064x" % g.qr, "%064x" % g.qm, "%064x" % g.qo, "%064x" % g.qk, g.a, g.b, g.c) for g in cs.gates]
    return head, gates


def _check(dumper, tmp_path, js, nvalues):
    r = _dump(dumper, tmp_path, js, nvalues)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    first = lines[0].split()
    src = [int(x) for x in lines[1].split()]
    gates = [tuple(t[:5]) + tuple(int(x) for x in t[5:]) for t in (ln.split() for ln in lines[2:])]
    head, want = _expect(js, nvalues)
    assert (int(first[0]), int(first[1]), src) == head
    assert gates == want
    return first[3]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_reference_fixtures(dumper, tmp_path, idx):
    js, vals = FIXTURES[idx]
    _check(dumper, tmp_path, js, len(vals))
    _check(dumper, tmp_path, js, 2)          # verify path: only the public inputs are passed (plonk.go:30)


def _random_acir(rng, nb_ops, nb_wit, pubs):
    def felt():
        return "%064x" % rng.choice([0, 1, o.R_MOD - 1, rng.randrange(o.R_MOD), rng.randrange(1 << 40)])

    def wit():
        return rng.randrange(0, nb_wit + 3)  # occasionally outside the value vector -> wire 0

    ops = []
    for _ in range(nb_ops):
        kind = rng.random()
        if kind < 0.75:
            mul = [[felt(), wit(), wit()] for _ in range(rng.choice([0, 1, 1, 2]))]
            lin = [[felt(), wit()] for _ in range(rng.choice([0, 1, 2, 3, 3, 4]))]
            ops.append({"Arithmetic": {"mul_terms": mul, "linear_combinations": lin, "q_c": felt()}})
        elif kind < 0.85:
            ops.append({"Directive": {"Invert": {"x": wit(), "result": wit()}}})
        elif kind < 0.92:
            ops.append({"Directive": {"ToRadix": {"a": {"mul_terms": [], "linear_combinations": [[felt(), 1]], "q_c": felt()},
                                                  "b": [1, 2, 3], "radix": 2, "note": 'br}ace ] in " string'}}})
        else:
            ops.append({"BlackBoxFuncCall": {"name": "RANGE", "inputs": [{"witness": wit(), "num_bits": 8}], "outputs": []}})
    return {"current_witness_index": nb_wit, "opcodes": ops, "public_inputs": pubs}


@pytest.mark.parametrize("seed", range(6))
def test_random_circuits_with_the_glue_quirks(dumper, tmp_path, seed):
    rng = random.Random(0xACE0 + seed)
    nb_wit = rng.randrange(3, 40)
    pubs = sorted(rng.sample(range(1, nb_wit + 1), rng.choice([0, 1, 1, 2, 3])))
    d = _random_acir(rng, rng.randrange(1, 60), nb_wit, pubs)
    # serde_json's compact form and a pretty-printed one with reordered keys must read the same
    compact = json.dumps(d, separators=(",", ":"))
    pretty = json.dumps({"public_inputs": d["public_inputs"], "extra": {"k": [1, {"z": "}"}]}, "opcodes": d["opcodes"],
                         "current_witness_index": d["current_witness_index"]}, indent=2)
    h1 = _check(dumper, tmp_path, compact, nb_wit)
    h2 = _check(dumper, tmp_path, pretty, nb_wit)
    assert h1 != h2                                      # the cache key is over the text
    _check(dumper, tmp_path, compact, max(1, nb_wit - 2))


@pytest.mark.parametrize("js,needle", [
    ('{"current_witness_index":1,"opcodes":[{"Foo":{}}],"public_inputs":[]}', "opcode"),
    ('{"current_witness_index":1,"opcodes":[{"Arithmetic":{"mul_terms":[],"q_c":"00"}}],"public_inputs":[]}', "Arithmetic"),
    ('{"current_witness_index":1,"opcodes":[{"Arithmetic":{"mul_terms":[],"linear_combinations":[["zz",1]],"q_c":"00"}}],"public_inputs":[]}', "hex"),
    ('{"current_witness_index":1,"opcodes":[{"Arithmetic":{"mul_terms":[[1,1,1]],"linear_combinations":[],"q_c":"00"}}],"public_inputs":[]}', "coefficient"),
    ('{"current_witness_index":1,"opcodes":[],"public_inputs":[]', "JSON"),
    ('{"opcodes":[],"public_inputs":[]}', "current witness"),
    ('{"current_witness_index":1,"opcodes":[{"BlackBoxFuncCall":{"name":"x"}}],"public_inputs":[]}', "BlackBoxFuncCall"),
])
def test_malformed_acir_is_fatal(dumper, tmp_path, js, needle):
    r = _dump(dumper, tmp_path, js, 1)
    assert r.returncode == 1 and needle in r.stderr, (r.returncode, r.stderr)


def test_large_circuit_reads_quickly(dumper, tmp_path):
    import time

    one, m1, zero = "%064x" % 1, "%064x" % (o.R_MOD - 1), "0" * 64
    n = 50000
    ops = ",".join('{"Arithmetic":{"mul_terms":[["%s",%d,%d]],"linear_combinations":[["%s",%d]],"q_c":"%s"}}' % (one, i + 1, i + 2, m1, i + 3, zero)
                   for i in range(n))
    js = '{"current_witness_index":%d,"opcodes":[%s],"public_inputs":[1]}' % (n + 2, ops)
    p = tmp_path / "big.json"
    p.write_text(js)
    t = time.time()
    r = subprocess.run([dumper, str(p), str(n + 2)], capture_output=True, text=True)
    dt = time.time() - t
    assert r.returncode == 0
    lines = r.stdout.splitlines()
    assert len(lines) == n + 2 and lines[0].split()[:2] == ["1", str(n + 1)]
    last = lines[-1].split()
    assert last[2] == one and last[3] == m1 and [int(x) for x in last[5:]] == [n - 1, n, n + 1]
    assert dt < 20
