"""ctypes binding of the C ABI declared in include/pikazoo_b200.h.

The product has no CPU path: if the CUDA library is missing this module raises, loudly.
"""

from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PIKAZOO_B200_LIB points at another build of the same library (tuning runs: profiles/time_kernels.py)
LIB_PATH = os.environ.get("PIKAZOO_B200_LIB") or os.path.join(HERE, "csrc", "libpikazoo_b200.so")

STATE_WORDS = 17
UNPACKED_WORDS = 53
OBS_WORDS = 35
NUM_STATS = 16

SERVE_CODES = {"winner": 0, "alternate": 1, "random": 2}
ACT_I32, ACT_I64, ACT_U8 = 0, 1, 2
REW_F32, REW_F64 = 0, 1
OBS_I32, OBS_I16, OBS_F32, OBS_F16, OBS_BF16, OBS_F64 = 0, 1, 2, 3, 4, 5
RINS_OFF, RINS_OUTER, RINS_INNER = 0, 1, 2
LAYOUT_ENV_MAJOR, LAYOUT_FEATURE_MAJOR, LAYOUT_ENV_MAJOR_SHARED = 0, 1, 2
ACTIONS_NOOP, ACTIONS_SYNTH = 0, 1
FLAG_NO_TABLES = 1
FLAG_NO_L2_HINTS = 2
FLAG_NO_PDL = 4
VERSION = 3

STAT_NAMES = (
    "calls", "episodes", "episode_frames", "p1_wins", "p2_wins", "p1_points", "p2_points", "resets",
    "bad_actions", "frozen", "truncated",
)


class PzConfig(ctypes.Structure):
    """struct pz_config (include/pikazoo_b200.h)."""

    _fields_ = [
        ("struct_bytes", ctypes.c_uint32),   # ABI handshake, filled by pz_config_init
        ("abi_version", ctypes.c_uint32),
        ("winning_score", ctypes.c_int32),
        ("serve", ctypes.c_int32),
        ("is_player1_computer", ctypes.c_int32),
        ("is_player2_computer", ctypes.c_int32),
        ("simplify_action", ctypes.c_int32),
        ("reward_by_ball_position", ctypes.c_int32),
        ("x_line", ctypes.c_int32),
        ("y_line", ctypes.c_int32),
        ("additional_reward", ctypes.c_double * 8),
        ("autoreset", ctypes.c_int32),
        ("action_dtype", ctypes.c_int32),
        ("reward_dtype", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("obs_dtype", ctypes.c_int32),
        ("normalize_observation", ctypes.c_int32),
        ("reward_in_normal_state", ctypes.c_int32),
        ("max_episode_frames", ctypes.c_int32),
        ("normal_state_reward", ctypes.c_double),
        ("obs_layout", ctypes.c_int32),
        ("obs_feature_rows", ctypes.c_int32),
    ]

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        # the handshake members, so that a PzConfig built field by field is accepted; load() has checked that
        # this declaration has the library's size
        self.struct_bytes = ctypes.sizeof(PzConfig)
        self.abi_version = VERSION


class PzEpisodeIo(ctypes.Structure):
    """struct pz_episode_io (include/pikazoo_b200.h)."""

    _fields_ = [
        ("episode_return_dev", ctypes.c_void_p),
        ("episode_length_dev", ctypes.c_void_p),
        ("truncated_dev", ctypes.c_void_p),
        ("status_dev", ctypes.c_void_p),
        ("seq_dev", ctypes.c_void_p),
        ("seq_value", ctypes.c_uint32),
    ]


class PikaLibraryError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load csrc/libpikazoo_b200.so (build it with `python pikazoo_b200/build.py`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PikaLibraryError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python pikazoo_b200/build.py` or __graft_entry__.build()). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, u64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32
    cfgp = ctypes.POINTER(PzConfig)
    L.pz_version.restype = ctypes.c_int
    L.pz_state_words.restype = ctypes.c_int
    L.pz_unpacked_words.restype = ctypes.c_int
    L.pz_state_bytes.argtypes = [i64]
    L.pz_state_bytes.restype = ctypes.c_size_t
    L.pz_strerror.argtypes = [ctypes.c_int]
    L.pz_strerror.restype = ctypes.c_char_p
    L.pz_config_bytes.restype = ctypes.c_size_t
    L.pz_config_init.argtypes = [cfgp, ctypes.c_size_t]
    L.pz_config_init.restype = ctypes.c_int
    L.pz_seed.argtypes = [vp, i64, u64, u64, vp]
    L.pz_seed_array.argtypes = [vp, i64, vp, vp]
    L.pz_reset.argtypes = [vp, i64, cfgp, vp, vp]
    L.pz_step.argtypes = [vp, i64, cfgp, vp, vp, vp, vp, vp, vp]
    epp = ctypes.POINTER(PzEpisodeIo)
    L.pz_reset_ex.argtypes = [vp, i64, cfgp, vp, epp, vp]
    L.pz_step_ex.argtypes = [vp, i64, cfgp, vp, vp, vp, vp, vp, epp, vp]
    L.pz_obs_elem_bytes.argtypes = [i32]
    L.pz_obs_elem_bytes.restype = ctypes.c_size_t
    L.pz_rollout.argtypes = [vp, i64, cfgp, i32, i32, u64, u64, u64, vp, vp, vp]
    L.pz_probe_write.argtypes = [vp, ctypes.c_size_t, i32, vp]
    L.pz_probe_write.restype = ctypes.c_int
    L.pz_export_state.argtypes = [vp, i64, vp, vp]
    L.pz_import_state.argtypes = [vp, i64, vp, vp]
    for name in ("pz_seed", "pz_seed_array", "pz_reset", "pz_reset_ex", "pz_step", "pz_step_ex", "pz_rollout",
                 "pz_export_state", "pz_import_state"):
        getattr(L, name).restype = ctypes.c_int
    L.pz_tables_prepare.argtypes = [vp]
    L.pz_tables_prepare.restype = ctypes.c_int
    L.pz_tables_bytes.restype = ctypes.c_size_t
    L.pz_tables_ready.restype = ctypes.c_int
    L.pz_tables_release.restype = None
    # host-buffer path
    L.pz_host_create.argtypes = [ctypes.POINTER(vp), i64, cfgp, u64, u64, i32]
    L.pz_host_create.restype = ctypes.c_int
    L.pz_host_reset.argtypes = [vp, vp]
    L.pz_host_reset.restype = ctypes.c_int
    L.pz_host_step.argtypes = [vp, vp, vp, vp, vp]
    L.pz_host_step.restype = ctypes.c_int
    L.pz_host_step_begin.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pz_host_step_begin.restype = ctypes.c_int
    L.pz_host_step_end.argtypes = [vp]
    L.pz_host_step_end.restype = ctypes.c_int
    L.pz_obs_player2_index.argtypes = [ctypes.c_int]
    L.pz_obs_player2_index.restype = ctypes.c_int
    L.pz_host_set_wire.argtypes = [vp, i32, i32]
    L.pz_host_set_wire.restype = ctypes.c_int
    L.pz_wire_expand.argtypes = [vp, vp, i64, i32, vp, i32, vp, vp]
    L.pz_wire_expand.restype = ctypes.c_int
    L.pz_host_stats.argtypes = [vp, vp]
    L.pz_host_stats.restype = ctypes.c_int
    L.pz_host_state_dev.argtypes = [vp]
    L.pz_host_state_dev.restype = vp
    L.pz_host_destroy.argtypes = [vp]
    L.pz_host_destroy.restype = None
    # caller side: fused MLP policy (pz_policy.cu)
    L.pz_policy_mlp_act.argtypes = [vp, i64, i64, i32, vp, i32, i32, vp, i32, i32, u64, u64, u64, vp, i32, i32, vp, vp]
    L.pz_policy_mlp_act.restype = ctypes.c_int
    L.pz_rollout_policy.argtypes = [vp, i64, cfgp, i32, vp, i32, i32, vp, i32, i32, u64, u64, u64, i32, vp, vp, vp, vp, vp]
    L.pz_rollout_policy.restype = ctypes.c_int
    L.pz_observe.argtypes = [vp, i64, cfgp, vp, vp]
    L.pz_observe.restype = ctypes.c_int
    L.pz_render.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp, vp]
    L.pz_render.restype = ctypes.c_int
    L.pz_policy_select.argtypes = [i32]
    L.pz_policy_select.restype = ctypes.c_int
    if (L.pz_version() != VERSION or L.pz_state_words() != STATE_WORDS or L.pz_unpacked_words() != UNPACKED_WORDS
            or L.pz_config_bytes() != ctypes.sizeof(PzConfig)):
        raise PikaLibraryError("libpikazoo_b200.so does not match this Python package (rebuild it)")
    _lib = L
    return L


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().pz_strerror(code).decode()
        raise PikaLibraryError(f"{what or 'pikazoo_b200'} failed: {msg} (code {code})")
