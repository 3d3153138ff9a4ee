"""In-tree build of libabyssb200.so (hand-written sm_100a CUDA behind the C ABI in include/abyss_b200.h).

    python -m abyss_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels with the
repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libabyssb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["abb_api.cu", "abb_assemble.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-shared", "-cudart", "static",
    "--expt-relaxed-constexpr", "-ccbin", "g++",
]


def _sources():
    out = []
    for d, _, fs in os.walk(CSRC):
        out += [os.path.join(d, f) for f in fs if f.endswith((".cu", ".cuh", ".h"))]
    for d, _, fs in os.walk(os.path.join(ROOT, "host")):
        out += [os.path.join(d, f) for f in fs if f.endswith((".cc", ".h"))]
    out.append(os.path.join(ROOT, "..", "include", "abyss_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC, *FLAGS, "--threads", "2", *(["-Xptxas", "-v"] if verbose else []), "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    build_cli()
    return LIB


def build_cli() -> None:
    """the C++ host programs (abyss-bloom-dbg, abyss-bloom) that link the C-ABI library"""
    host = os.path.join(ROOT, "host")
    for exe, src in (("abyss-bloom-dbg", "abyss_bloom_dbg.cc"), ("abyss-bloom", "abyss_bloom.cc")):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-o", os.path.join(LIBDIR, exe), os.path.join(host, src),
               "-L" + LIBDIR, "-labyssb200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
