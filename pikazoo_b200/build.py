"""Build the sm_100a shared library (csrc/libpikazoo_b200.so) in-tree with nvcc.

    python pikazoo_b200/build.py [--force] [--verbose] [--out variants/NAME.so]

--out builds a tuning variant (with PZ_NVCC_FLAGS=-D...) beside the product library instead of replacing it;
PIKAZOO_B200_LIB=<that path> makes the package load it.

nvcc cross-compiles without a GPU; the built .so is git-ignored and travels with the tree.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(CSRC, "libpikazoo_b200.so")
SOURCES = ["pz_kernels.cu", "pz_host.cu", "pz_policy.cu", "pz_policy_tc.cu", "pz_rollout_policy.cu", "pz_render.cu", "pz_wire.cpp", "pz_step_ai0.cu", "pz_step_ai1.cu", "pz_step_ai2.cu", "pz_step_ai3.cu"]
HEADERS = ["pz_state.cuh", "pz_rng.cuh", "pz_physics.cuh", "pz_kernels.cuh", "pz_device.cuh", "pz_policy.cuh", "pz_tcgen05.cuh", "pz_step_inst.inc"]


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


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    if out is None and not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
             "-I", INCLUDE]
    flags += os.environ.get("PZ_NVCC_FLAGS", "").split()  # e.g. -DPZ_ROLLOUT_MIN_CTAS=4 for tuning runs
    if verbose:
        flags += ["-Xptxas", "-v"]
    lib_path = LIB_PATH if out is None else os.path.abspath(out)
    objdir = os.path.join(CSRC, "build") if out is None else os.path.splitext(lib_path)[0] + "_obj"
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    # the translation units are independent (no relocatable device code): compile them in parallel
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(f"---- {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    # (the link step gets the architecture too: without it nvcc adds an empty device-link image for its default sm_52)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib_path]
                   + [obj for _, obj, _ in results], check=True)
    return lib_path


if __name__ == "__main__":
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=out_path))
