"""CPU: the C-ABI library builds, loads, and exports every symbol include/*.h declares.
No compute calls (there is no GPU here)."""

import ctypes
import glob
import os
import re

from tests.conftest import ROOT


def _declared_symbols():
    syms = set()
    for path in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        syms |= set(re.findall(r"\b(pz_[a-z0-9_]+)\s*\(", text))
    return syms


def test_header_declares_the_path():
    syms = _declared_symbols()
    assert {"pz_seed", "pz_reset", "pz_step", "pz_rollout", "pz_export_state", "pz_import_state",
            "pz_host_create", "pz_host_step", "pz_strerror", "pz_step_ex", "pz_reset_ex", "pz_obs_elem_bytes"} <= syms


def test_library_exports_every_declared_symbol(cuda_lib):
    for name in sorted(_declared_symbols()):
        assert hasattr(cuda_lib, name), f"{name} declared in include/ but not exported"


def test_constants_and_errors(cuda_lib):
    assert cuda_lib.pz_version() == 2
    assert cuda_lib.pz_state_words() == 17
    assert cuda_lib.pz_unpacked_words() == 53
    assert cuda_lib.pz_state_bytes(1000) == 1000 * 17 * 4
    assert b"bad config" in cuda_lib.pz_strerror(-2)
    assert cuda_lib.pz_strerror(0) == b"success"


def test_argument_validation_without_gpu(cuda_lib):
    # argument checks happen before any CUDA call, so they are testable on CPU
    from pikazoo_b200 import make_config

    cfg = make_config()
    assert cuda_lib.pz_step(None, 4, ctypes.byref(cfg), None, None, None, None, None, None) == -1
    assert cuda_lib.pz_seed(None, 4, 0, 0, None) == -1
    assert cuda_lib.pz_rollout(ctypes.c_void_p(16), 4, ctypes.byref(cfg), 0, 0, 0, 0, 0, None, None, None) == -1
    cfg.winning_score = 5000
    assert cuda_lib.pz_reset(ctypes.c_void_p(16), 4, ctypes.byref(cfg), None, None) == -2
    cfg.winning_score = 15
    assert cuda_lib.pz_reset(ctypes.c_void_p(8), 4, ctypes.byref(cfg), None, None) == -3  # misaligned
    assert cuda_lib.pz_reset(ctypes.c_void_p(16), 0, ctypes.byref(cfg), None, None) == 0  # empty batch
    cfg.obs_dtype = 9
    assert cuda_lib.pz_reset(ctypes.c_void_p(16), 4, ctypes.byref(cfg), None, None) == -2
    cfg.obs_dtype, cfg.normalize_observation = 0, 1  # NormalizeObservation needs a float dtype
    assert cuda_lib.pz_step_ex(ctypes.c_void_p(16), 4, ctypes.byref(cfg), ctypes.c_void_p(16), None, None, None, None,
                               None, None) == -2
    assert [cuda_lib.pz_obs_elem_bytes(k) for k in range(7)] == [4, 2, 4, 2, 2, 8, 0]
    cfg.normalize_observation, cfg.obs_layout, cfg.obs_feature_rows = 0, 1, 20  # fewer rows than an observation
    assert cuda_lib.pz_reset(ctypes.c_void_p(16), 4, ctypes.byref(cfg), None, None) == -2


def test_product_does_not_import_oracle():
    # the product path must never route through the oracle
    pkg = os.path.join(ROOT, "pikazoo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pika_oracle" not in text.replace("oracle/pika_oracle.h", ""), f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_config_struct_layout_matches_header(cuda_lib):
    from pikazoo_b200._lib import PzConfig

    c = PzConfig()
    cuda_lib.pz_default_config(ctypes.byref(c))
    assert (c.winning_score, c.serve, c.x_line, c.y_line, c.autoreset) == (15, 0, 216, 176, 1)
    assert ctypes.sizeof(PzConfig) == 8 * 4 + 8 * 8 + 8 * 4 + 8 + 2 * 4 and PzConfig.flags.offset == 108
    assert PzConfig.normal_state_reward.offset == 128 and c.obs_dtype == 0 and c.max_episode_frames == 0
