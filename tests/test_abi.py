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
    assert cuda_lib.pz_version() == 3
    assert cuda_lib.pz_state_words() == 17
    assert cuda_lib.pz_unpacked_words() == 53
    assert cuda_lib.pz_state_bytes(1000) == 1000 * 17 * 4
    assert b"bad config" in cuda_lib.pz_strerror(-2)
    assert cuda_lib.pz_strerror(0) == b"success"
    assert b"struct_bytes" in cuda_lib.pz_strerror(-5)


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
    from pikazoo_b200._lib import VERSION, PzConfig

    c = PzConfig()
    assert cuda_lib.pz_config_init(ctypes.byref(c), ctypes.sizeof(c)) == 0
    assert (c.winning_score, c.serve, c.x_line, c.y_line, c.autoreset) == (15, 0, 216, 176, 1)
    assert (c.struct_bytes, c.abi_version) == (ctypes.sizeof(PzConfig), VERSION)
    assert ctypes.sizeof(PzConfig) == cuda_lib.pz_config_bytes() == 2 * 4 + 8 * 4 + 8 * 8 + 8 * 4 + 8 + 2 * 4
    assert PzConfig.flags.offset == 116 and PzConfig.additional_reward.offset == 40
    assert PzConfig.normal_state_reward.offset == 136 and c.obs_dtype == 0 and c.max_episode_frames == 0


def test_abi_handshake_refuses_a_stale_struct(cuda_lib):
    """A binding that declares another revision of struct pz_config (the documented stub was 8 bytes
    short) is refused with PZ_E_ABI before anything else is read, and pz_config_init writes nothing to it."""
    from pikazoo_b200 import make_config

    buf = (ctypes.c_ubyte * 256)(*([0xAB] * 256))
    cfgp = ctypes.cast(buf, ctypes.POINTER(type(make_config())))
    for wrong in (cuda_lib.pz_config_bytes() - 8, cuda_lib.pz_config_bytes() + 8, 0):
        assert cuda_lib.pz_config_init(cfgp, wrong) == -5
        assert bytes(buf) == b"\xab" * 256
    cfg = make_config()
    for field, bad in (("struct_bytes", cfg.struct_bytes - 8), ("abi_version", cfg.abi_version - 1)):
        c = make_config()
        setattr(c, field, bad)
        assert cuda_lib.pz_reset(ctypes.c_void_p(16), 4, ctypes.byref(c), None, None) == -5
        assert cuda_lib.pz_step(ctypes.c_void_p(16), 4, ctypes.byref(c), ctypes.c_void_p(16), None, None, None, None,
                                None) == -5
        assert cuda_lib.pz_rollout(ctypes.c_void_p(16), 4, ctypes.byref(c), 1, 0, 0, 0, 0, None, None, None) == -5
        ctx = ctypes.c_void_p()
        assert cuda_lib.pz_host_create(ctypes.byref(ctx), 4, ctypes.byref(c), 0, 0, 1) == -5 and not ctx.value


def integration_stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.search(r"```python\n(# pikazoo/env/b200_backend\.py.*?)```", text, flags=re.S).group(1)


def test_integration_md_stub_declares_this_librarys_struct(cuda_lib):
    """The ctypes stub INTEGRATION.md shows is parsed out of the document: its _Cfg must be the library's struct
    (size and every member's offset = the package's own binding)."""
    from pikazoo_b200 import _lib

    block = integration_stub_source().replace('ctypes.CDLL("libpikazoo_b200.so")', f'ctypes.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)  # noqa: S102 - our own document; no compute at import
    stub = ns["_Cfg"]
    assert ctypes.sizeof(stub) == cuda_lib.pz_config_bytes()
    assert [(n, getattr(stub, n).offset) for n, _ in stub._fields_] == \
        [(n, getattr(_lib.PzConfig, n).offset) for n, _ in _lib.PzConfig._fields_]
