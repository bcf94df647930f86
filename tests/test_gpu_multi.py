"""GPU, >= 2 devices: the N > 1 path under its real launcher — torchrun, one process per GPU, NCCL — instead of an
emulation on one device (tests/test_gpu_scale.py::test_sharding_is_invisible). Skipped on single-GPU boxes; run it
with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""

import os
import subprocess
import sys

import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_under_torchrun(cuda_lib):
    world = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29741", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTI_GPU_CHECK OK" in r.stdout
