"""Build the sm_100a shared library (csrc/libpikazoo_b200.so) in-tree with nvcc.

    python pika-zoo_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the built .so is git-ignored and travels with the tree.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(CSRC, "libpikazoo_b200.so")
SOURCES = ["pz_kernels.cu", "pz_host.cu"]
HEADERS = ["pz_state.cuh", "pz_rng.cuh", "pz_physics.cuh", "pz_kernels.cuh"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "pikazoo_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared",
        "-I", INCLUDE,
        "-o", LIB_PATH,
    ] + [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
