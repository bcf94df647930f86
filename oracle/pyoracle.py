"""TEST INFRASTRUCTURE — ctypes binding of the C oracle (oracle/pika_oracle.c).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs. The product package never imports it.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpika_oracle.so")

ENV_WORDS = 53
SERVE_CODES = {"winner": 0, "alternate": 1, "random": 2}


class PkConfig(ctypes.Structure):
    _fields_ = [
        ("winning_score", ctypes.c_int32),
        ("serve", ctypes.c_int32),
        ("is_player1_computer", ctypes.c_int32),
        ("is_player2_computer", ctypes.c_int32),
        ("simplify_action", ctypes.c_int32),
        ("reward_by_ball_position", ctypes.c_int32),
        ("x_line", ctypes.c_int32),
        ("y_line", ctypes.c_int32),
        ("additional_reward", ctypes.c_double * 8),
        ("reward_in_normal_state", ctypes.c_int32),
        ("max_episode_frames", ctypes.c_int32),
        ("normal_state_reward", ctypes.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms). Returns the .so path."""
    src = os.path.join(_HERE, "pika_oracle.c")
    hdr = os.path.join(_HERE, "pika_oracle.h")
    if (
        force
        or not os.path.exists(LIB_PATH)
        or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, u64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32
        cfgp = ctypes.POINTER(PkConfig)
        L.pk_env_words.restype = ctypes.c_int
        L.pk_pcg64_seed.argtypes = [u64, vp, vp]
        L.pk_integers.argtypes = [vp, u32]
        L.pk_integers.restype = i32
        L.pk_init.argtypes = [vp, u64]
        L.pk_reset.argtypes = [vp, cfgp, vp]
        L.pk_step.argtypes = [vp, cfgp, i32, i32, vp, vp, vp]
        L.pk_step.restype = ctypes.c_int
        L.pk_vec_step.argtypes = [vp, i64, cfgp, vp, vp, vp, vp, ctypes.c_int]
        L.pk_vec_step.restype = ctypes.c_int
        L.pk_vec_step_ex.argtypes = [vp, i64, cfgp, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp]
        L.pk_vec_step_ex.restype = ctypes.c_int
        L.pk_vec_reset_ex.argtypes = [vp, i64, cfgp, vp, vp, vp]
        L.pk_normalize_obs.argtypes = [i64, vp, vp]
        L.pk_vec_obs.argtypes = [vp, i64, vp]
        L.pk_vec_init.argtypes = [vp, i64, u64]
        L.pk_vec_reset.argtypes = [vp, i64, cfgp, vp]
        L.pk_synth_action.argtypes = [u64, u64, u64, ctypes.c_int, u32]
        L.pk_synth_action.restype = i32
        L.pk_vec_rollout.argtypes = [vp, i64, cfgp, ctypes.c_int, ctypes.c_int, u64, u64, u64, vp]
        L.pk_vec_rollout.restype = i64
        L.pk_simulate_many.argtypes = [i64, vp, ctypes.c_int, vp]
        L.pk_simulate_many.restype = None
        assert L.pk_env_words() == ENV_WORDS
        _lib = L
    return _lib


def make_config(
    winning_score=15,
    serve="winner",
    is_player1_computer=False,
    is_player2_computer=False,
    simplify_action=False,
    reward_by_ball_position=None,
    reward_in_normal_state=None,
    normal_state_first=False,
    max_episode_frames=0,
    normalize_observation=False,        # handled by OracleVecEnv.normalized_obs / convert_obs
    record_episode_statistics=False,    # OracleVecEnv always keeps episode_return / episode_length
) -> PkConfig:
    c = PkConfig()
    c.winning_score = int(winning_score)
    c.serve = SERVE_CODES[serve]
    c.is_player1_computer = int(bool(is_player1_computer))
    c.is_player2_computer = int(bool(is_player2_computer))
    c.simplify_action = int(bool(simplify_action))
    c.x_line, c.y_line = 216, 176
    if reward_by_ball_position is not None:
        add, x_line, y_line = reward_by_ball_position
        assert len(add) == 8
        c.reward_by_ball_position = 1
        c.x_line, c.y_line = int(x_line), int(y_line)
        for k in range(8):
            c.additional_reward[k] = float(add[k])
    if reward_in_normal_state is not None:
        c.reward_in_normal_state = 2 if normal_state_first else 1
        c.normal_state_reward = float(reward_in_normal_state)
    c.max_episode_frames = int(max_episode_frames)
    return c


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleVecEnv:
    """N independent oracle envs with the product's batched NEXT-STEP auto-reset semantics."""

    def __init__(self, num_envs, seed=0, autoreset=True, seeds=None, **cfg):
        self.n = int(num_envs)
        self.cfg = make_config(**cfg)
        self.autoreset = bool(autoreset)
        self.state = np.zeros((self.n, ENV_WORDS), dtype=np.int32)
        self.obs = np.zeros((self.n, 2, 35), dtype=np.int32)
        self.reward = np.zeros((self.n, 2), dtype=np.float64)
        self.done = np.zeros((self.n,), dtype=np.uint8)
        self.truncated = np.zeros((self.n,), dtype=np.uint8)
        self.episode_return = np.zeros((self.n, 2), dtype=np.float64)
        self.episode_length = np.zeros((self.n,), dtype=np.int32)
        lib().pk_vec_init(_p(self.state), self.n, int(seed))
        if seeds is not None:  # explicit per-env seeds (a strided sample of a larger batch)
            assert len(seeds) == self.n
            for j, sd in enumerate(seeds):
                lib().pk_init(_p(self.state[j]), int(sd))

    def reset(self):
        lib().pk_vec_reset_ex(_p(self.state), self.n, ctypes.byref(self.cfg), _p(self.obs), _p(self.episode_return),
                              _p(self.episode_length))
        return self.obs

    def step(self, actions):
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, 2)
        rc = lib().pk_vec_step_ex(
            _p(self.state), self.n, ctypes.byref(self.cfg), None if a is None else _p(a), _p(self.obs),
            _p(self.reward), _p(self.done), int(self.autoreset), _p(self.episode_return), _p(self.episode_length),
            _p(self.truncated),
        )
        if rc != 0:
            raise IndexError("action out of range")
        return self.obs, self.reward, self.done

    def current_obs(self):
        """_get_obs of every env from the current state (e.g. after a rollout)."""
        lib().pk_vec_obs(_p(self.state), self.n, _p(self.obs))
        return self.obs

    def normalized_obs(self, dtype=np.float64):
        """NormalizeObservation output for the current obs: float64 as the reference wrapper produces
        it; float32 = astype(float32) of that; float16 = astype(float16) of the float32; "bfloat16" =
        the float32 rounded to nearest-even on its upper 16 bits, returned as uint16 bit patterns."""
        return convert_obs(self.obs, dtype, normalize=True)

    def raw_obs_bf16(self):
        """(float)value rounded to bfloat16 (PZ_OBS_BF16 without NormalizeObservation), as uint16 bit patterns."""
        return convert_obs(self.obs, "bfloat16", normalize=False)

    def rollout(self, K, action_mode=0, action_seed=0, first_env=0, frame0=0, stats=None):
        if stats is None:
            stats = np.zeros(9, dtype=np.int64)
        lib().pk_vec_rollout(
            _p(self.state), self.n, ctypes.byref(self.cfg), int(K), int(action_mode), int(action_seed),
            int(first_env), int(frame0), _p(stats),
        )
        return stats


def pcg64_seed(seed: int):
    st = np.zeros(4, dtype=np.uint32)
    inc = np.zeros(4, dtype=np.uint32)
    lib().pk_pcg64_seed(int(seed), _p(st), _p(inc))
    return st, inc


def synth_action(action_seed, global_env, frame, agent, n_actions=18) -> int:
    return int(lib().pk_synth_action(int(action_seed), int(global_env), int(frame), int(agent), int(n_actions)))


def simulate_many(xyv: np.ndarray, power: bool) -> np.ndarray:
    """Landing x of the reference's trajectory loops from starts xyv[n, 4] = (x, y, xv, yv)."""
    q = np.ascontiguousarray(xyv, dtype=np.int32).reshape(-1, 4)
    out = np.zeros(len(q), dtype=np.int32)
    lib().pk_simulate_many(len(q), _p(q), int(bool(power)), _p(out))
    return out


def normalize_obs(obs: np.ndarray) -> np.ndarray:
    o = np.ascontiguousarray(obs, dtype=np.int32).reshape(-1, 70)
    out = np.zeros(o.shape, dtype=np.float64)
    lib().pk_normalize_obs(len(o), _p(o), _p(out))
    return out.reshape(np.shape(obs))


def float32_to_bfloat16_bits(f: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even of float32 to bfloat16, as uint16 bit patterns (no NaNs occur here)."""
    u = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def convert_obs(obs: np.ndarray, dtype, normalize: bool):
    """The product's observation dtypes restated with numpy (include/pikazoo_b200.h PZ_OBS_*)."""
    if dtype in (np.int32, np.int16):
        assert not normalize
        return np.asarray(obs).astype(dtype)
    f64 = normalize_obs(obs) if normalize else np.asarray(obs).astype(np.float64)
    if dtype == np.float64:
        return f64
    f32 = f64.astype(np.float32)
    if dtype == np.float32:
        return f32
    if dtype == np.float16:
        return f32.astype(np.float16)
    if dtype == "bfloat16":
        return float32_to_bfloat16_bits(f32)
    raise ValueError(dtype)
