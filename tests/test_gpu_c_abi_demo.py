"""GPU: the C ABI used from plain C (examples/c_abi_demo.c: no Python, no torch in the process) — the
program is compiled with gcc against include/pikazoo_b200.h, run, and its checksums are compared with
the oracle's for the same seeds and actions."""

import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _fnv1a(h, data: bytes) -> int:
    """examples/c_abi_demo.c fold(): h * 0x100000001B3 + sum(byte[i] * (i + 1)) mod 2^64"""
    a = np.frombuffer(data, dtype=np.uint8).astype(np.uint64)
    with np.errstate(over="ignore"):
        s = int((a * np.arange(1, len(a) + 1, dtype=np.uint64)).sum(dtype=np.uint64))
    return (h * 0x100000001B3 + s) & 0xFFFFFFFFFFFFFFFF


def test_c_program_against_oracle(cuda_lib, tmp_path):
    from pikazoo_b200 import _lib

    exe = str(tmp_path / "c_abi_demo")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", os.path.join(ROOT, "examples", "c_abi_demo.c"), "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(cuda, "include"), "-L", libdir, "-lpikazoo_b200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    n, steps, seed = 640 + 9, 150, 3
    out = subprocess.run([exe, str(n), str(steps), str(seed)], check=True, capture_output=True, text=True).stdout
    dev_line, host_line = out.strip().splitlines()

    cfg = dict(is_player1_computer=True, is_player2_computer=True, winning_score=3, serve="random")
    orc = po.OracleVecEnv(n, seed=seed, **cfg)
    orc.reset()
    episodes = resets = 0
    for t in range(steps):
        was_over = orc.state[:, 40].copy()
        orc.step(None)
        resets += int(was_over.sum())
        episodes += int((orc.done != 0).sum() - 0)
    h = _fnv1a(0xCBF29CE484222325, orc.state.tobytes())
    assert dev_line.split()[:4] == ["device", str(n), str(steps), f"{h:016x}"], dev_line
    assert f"resets={resets}" in dev_line

    hcfg = dict(simplify_action=True, winning_score=2)
    orc = po.OracleVecEnv(n, seed=seed, **hcfg)
    h = _fnv1a(0xCBF29CE484222325, orc.reset().tobytes())
    for t in range(steps):
        obs, rew, done = orc.step(synth_actions_numpy(7, 0, n, t, 13))
        h = _fnv1a(h, obs.tobytes())
        h = _fnv1a(h, rew.astype(np.float32).tobytes())
        h = _fnv1a(h, done.tobytes())
    assert host_line.split() == ["host", str(n), str(steps), f"{h:016x}"], host_line
