"""In-tree build of libb200zk.so (the sm_100a CUDA library behind include/b200zk.h) and of
libgnark_backend_b200.so (the C++ stand-in for the reference's cgo exports, include/gnark_backend_ffi.h).

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels with the repo
snapshot to the GPU box.  Usage: python -m noir_backend_using_gnark_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libb200zk.so"
FFI_LIB = LIBDIR / "libgnark_backend_b200.so"
FFI_SRC = CSRC / "ffi" / "gnark_backend_ffi.cpp"
SOURCES = ["capi.cu", "ntt.cu", "msm.cu", "srs.cu", "microbench.cu", "plonk.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "550",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    inc = PKG.parent / "include"
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((CSRC / "ffi").glob("*"))
                    + [inc / "b200zk.h", inc / "gnark_backend_ffi.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / "libb200zk.stamp"
    digest = _digest()
    if not force and LIB.exists() and FFI_LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    # host-only C++ on top of the C ABI; found next to libb200zk.so through $ORIGIN
    cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", str(FFI_LIB), str(FFI_SRC),
           "-L" + str(LIBDIR), "-lb200zk", "-pthread", "-Wl,-rpath,$ORIGIN",
           "-Wl,--exclude-libs,ALL"]  # keep the statically linked parts of the C++ runtime out of the export table
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for %s:\n%s\n%s" % (FFI_SRC.name, r.stdout, r.stderr))
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
