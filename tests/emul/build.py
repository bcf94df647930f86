"""TEST INFRASTRUCTURE: builds tests/emul/pz_emul.cpp (the device headers compiled for the host)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libpz_emul.so")
DEPS = [os.path.join(HERE, "pz_emul.cpp")] + [
    os.path.join(ROOT, "pikazoo_b200", "csrc", f) for f in ("pz_state.cuh", "pz_rng.cuh", "pz_physics.cuh", "pz_policy.cuh")]


def cuda_include():
    for c in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if c and os.path.exists(os.path.join(c, "include", "cuda_runtime.h")):
            return os.path.join(c, "include")
    return None


def build():
    inc = cuda_include()
    if inc is None:
        return None
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in DEPS):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", inc, "-I", os.path.join(ROOT, "include"), "-o", LIB, DEPS[0]], check=True)
    return LIB
