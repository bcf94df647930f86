"""Batched device-tensor entry point: N independent Pikachu-Volleyball envs on one B200.

`PikaVecEnv` keeps the reference constructor arguments (pikazoo/env/pikazoo_env.py:79-86:
winning_score, serve, is_player1_computer, is_player2_computer) and fuses the two wrappers
north_star names (SimplifyAction, simplify_action.py:7-28; RewardByBallPosition,
reward_by_ball_position.py:6-31) into the step kernel. All tensors live on the CUDA device;
every call is asynchronous on the current torch stream. PyTorch is used for device memory
and streams only — the simulation is the hand-written sm_100a library behind the C ABI
(include/pikazoo_b200.h). There is no CPU fallback.

Semantics of one `step(actions)` call per env (NEXT-STEP auto-reset):
  * env not terminated: exactly one reference `env.step({"player_1": a1, "player_2": a2})`;
  * env terminated by an earlier call and autoreset=True: exactly one reference `env.reset()`
    on the same object (carry-over semantics), obs = reset obs, reward 0, done False, action ignored;
  * env terminated and autoreset=False: no-op, obs re-emitted, reward 0, done True.
Env i's random stream is numpy `Generator(PCG64(seed + first_env + i))` (protocol S0): seed s + 1 is the
batch of seed s shifted by one env, so independent replicates must space their seeds by at least the total
number of envs (e.g. seed = replicate * total_envs).
"""

from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

_ACT_DTYPES = {torch.int32: _lib.ACT_I32, torch.int64: _lib.ACT_I64, torch.uint8: _lib.ACT_U8}
_REW_DTYPES = {torch.float32: _lib.REW_F32, torch.float64: _lib.REW_F64}
_OBS_DTYPES = {torch.int32: _lib.OBS_I32, torch.int16: _lib.OBS_I16, torch.float32: _lib.OBS_F32,
               torch.float16: _lib.OBS_F16, torch.bfloat16: _lib.OBS_BF16, torch.float64: _lib.OBS_F64}

# torch.cuda.current_stream() / current_device() walk several Python layers (3-5 us per call — more than the launch
# they precede); the raw accessors behind them are one C call each
try:
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _raw_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover - other torch builds
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream

    _raw_device = torch.cuda.current_device

_LAYOUTS = {"env_major": _lib.LAYOUT_ENV_MAJOR, "feature_major": _lib.LAYOUT_FEATURE_MAJOR,
            "shared": _lib.LAYOUT_ENV_MAJOR_SHARED}

AGENTS = ("player_1", "player_2")
# obs_layout="shared": player_2's observation = row[PLAYER2_INDEX] (the two player blocks swapped, pikazoo_env.py:585-586)
PLAYER2_INDEX = tuple(list(range(13, 26)) + list(range(0, 13)) + list(range(26, 35)))


def make_config(
    winning_score: int = 15,
    serve: str = "winner",
    is_player1_computer: bool = False,
    is_player2_computer: bool = False,
    simplify_action: bool = False,
    reward_by_ball_position: Optional[Tuple[Sequence[float], int, int]] = None,
    autoreset: bool = True,
    action_dtype: torch.dtype = torch.int32,
    reward_dtype: torch.dtype = torch.float32,
    landing_tables: bool = True,
    obs_dtype: torch.dtype = torch.int32,
    normalize_observation: bool = False,
    reward_in_normal_state: Optional[float] = None,
    normal_state_first: bool = False,
    max_episode_frames: int = 0,
    l2_hints: bool = True,
    obs_layout: str = "env_major",
    obs_feature_rows: int = 35,
    pdl: bool = True,
) -> _lib.PzConfig:
    assert serve in ("winner", "alternate", "random")  # pikazoo_env.py:104
    if not 1 <= int(winning_score) <= 1023:
        raise ValueError("winning_score must be in [1, 1023] (10-bit packed score field)")
    c = _lib.PzConfig()
    c.winning_score = int(winning_score)
    c.serve = _lib.SERVE_CODES[serve]
    c.is_player1_computer = int(bool(is_player1_computer))
    c.is_player2_computer = int(bool(is_player2_computer))
    c.simplify_action = int(bool(simplify_action))
    c.x_line, c.y_line = 216, 176
    if reward_by_ball_position is not None:
        add, x_line, y_line = reward_by_ball_position
        assert len(add) == 8  # reward_by_ball_position.py:15
        c.reward_by_ball_position = 1
        c.x_line, c.y_line = int(x_line), int(y_line)
        for k in range(8):
            c.additional_reward[k] = float(add[k])
    c.autoreset = int(bool(autoreset))
    c.action_dtype = _ACT_DTYPES[action_dtype]
    c.reward_dtype = _REW_DTYPES[reward_dtype]
    c.flags = ((0 if landing_tables else _lib.FLAG_NO_TABLES) | (0 if l2_hints else _lib.FLAG_NO_L2_HINTS)
               | (0 if pdl else _lib.FLAG_NO_PDL))
    if obs_dtype not in _OBS_DTYPES:
        raise TypeError(f"obs_dtype must be one of {sorted(str(k) for k in _OBS_DTYPES)}")
    c.obs_dtype = _OBS_DTYPES[obs_dtype]
    if normalize_observation and not obs_dtype.is_floating_point:
        raise TypeError("normalize_observation needs a floating-point obs_dtype "
                        "(the reference wrapper yields float64; its declared space is float32)")
    c.normalize_observation = int(bool(normalize_observation))
    if reward_in_normal_state is not None:  # reward_in_normal_state.py:7-9
        c.reward_in_normal_state = _lib.RINS_INNER if normal_state_first else _lib.RINS_OUTER
        c.normal_state_reward = float(reward_in_normal_state)
    if int(max_episode_frames) < 0:
        raise ValueError("max_episode_frames must be >= 0 (0 = never truncate)")
    c.max_episode_frames = int(max_episode_frames)
    if obs_layout not in _LAYOUTS:
        raise ValueError("obs_layout must be 'env_major', 'feature_major' or 'shared'")
    if obs_layout == "shared" and obs_dtype not in (torch.int32, torch.int16):
        raise TypeError("obs_layout='shared' carries the integer observation: obs_dtype must be int32 or int16")
    c.obs_layout = _LAYOUTS[obs_layout]
    if int(obs_feature_rows) < _lib.OBS_WORDS:
        raise ValueError("obs_feature_rows must be >= 35")
    c.obs_feature_rows = int(obs_feature_rows)
    return c


class PikaVecEnv:
    """N envs stepped by one kernel launch; observations [N, 2, 35] int32 on the device."""

    possible_agents = list(AGENTS)
    num_agents = 2
    obs_dim = _lib.OBS_WORDS

    def __init__(
        self,
        num_envs: int,
        device: "torch.device | str | int" = "cuda",
        seed: int = 0,
        winning_score: int = 15,
        serve: str = "winner",
        is_player1_computer: bool = False,
        is_player2_computer: bool = False,
        simplify_action: bool = False,
        reward_by_ball_position: Optional[Tuple[Sequence[float], int, int]] = None,
        autoreset: bool = True,
        action_dtype: torch.dtype = torch.int32,
        reward_dtype: torch.dtype = torch.float32,
        first_env: int = 0,
        track_stats: bool = True,
        landing_tables: "bool | str" = "auto",
        obs_dtype: torch.dtype = torch.int32,
        normalize_observation: bool = False,
        reward_in_normal_state: Optional[float] = None,
        normal_state_first: bool = False,
        max_episode_frames: int = 0,
        record_episode_statistics: bool = False,
        l2_hints: bool = True,
        obs_layout: str = "env_major",
        obs_feature_rows: int = 35,
        pdl: bool = True,
        host_mapped: bool = False,
        status: bool = False,
    ):
        """Beyond the reference's constructor arguments:

        obs_layout: "env_major" — obs [N, 2, 35], one row per env and agent like the reference — or
          "feature_major" — obs [2, obs_feature_rows, N]: every feature a contiguous vector over envs, the
          layout a device-side policy consumes without a transposing pass (GEMM operand with leading
          dimension N). Rows 35.. of each agent (obs_feature_rows=40 pads K to a multiple of 8 for
          tensor-core GEMMs) stay zero.
          "shared" — obs [N, 35] (int32 / int16): player_1's row only; player_2's observation is
          `obs[:, PLAYER2_INDEX]` (the two player blocks swapped): a half (a quarter as int16) of the bytes.
        status: also emit `self.status` uint8 [N] = (player_1's base reward + 1) | terminated << 2 | truncated << 3.
        obs_dtype: torch.int32 (the reference's declared dtype), int16 (same integers, half the bytes),
          or float32 / float16 / bfloat16 / float64 — `(float)value`, or with normalize_observation=True
          the NormalizeObservation wrapper's output (normalize_observation.py:18-32) computed in the
          kernel; float64 is bit-identical to the reference wrapper, float32 is float32() of it.
        reward_in_normal_state: fuses RewardInNormalState(env, reward) (reward_in_normal_state.py:10-15),
          applied outside RewardByBallPosition, or inside it with normal_state_first=True.
        max_episode_frames: truncate episodes that reach this many step() calls (0 = never, like the
          reference); `self.truncated` [N] bool reports it and the next call resets the env.
        host_mapped: every per-call buffer (obs, reward, done, status, episode statistics) is pinned host memory,
          which the device addresses directly (unified addressing): the kernels read and write it over PCIe, so a
          host-driven step of a tiny batch is one launch and one stream synchronisation, no copies — the
          single-env facade's mode. `step()` then takes a pinned CPU action tensor and the returned tensors are
          CPU tensors, valid after a synchronisation of the stream. Pointless for large batches.
        record_episode_statistics: fuses RecordEpisodeStatistics (record_episode_statistics.py:17-40):
          `self.episode_return` [N, 2] float64 and `self.episode_length` [N] int32 are valid for env i
          on the call where terminated[i] (or truncated[i]) is set.
        """
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PikaLibraryError("PikaVecEnv runs on CUDA devices only (there is no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._dev_index = self.device.index
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.seed = int(seed)
        self.first_env = int(first_env)
        # Computer players can read their trajectory simulations from memoised per-device tables
        # (1.9 GB of HBM, built once on first use; results identical). "auto": only for batches
        # large enough for the tables to pay for themselves.
        if landing_tables == "auto":
            landing_tables = self.num_envs >= 4096
        self.landing_tables = bool(landing_tables)
        self._kw = dict(
            winning_score=winning_score, serve=serve, is_player1_computer=is_player1_computer,
            is_player2_computer=is_player2_computer, simplify_action=simplify_action,
            reward_by_ball_position=reward_by_ball_position, autoreset=autoreset,
            action_dtype=action_dtype, reward_dtype=reward_dtype, landing_tables=self.landing_tables,
            obs_dtype=obs_dtype, normalize_observation=normalize_observation,
            reward_in_normal_state=reward_in_normal_state, normal_state_first=normal_state_first,
            max_episode_frames=max_episode_frames, l2_hints=l2_hints, obs_layout=obs_layout,
            obs_feature_rows=obs_feature_rows, pdl=pdl,
        )
        self.cfg = make_config(**self._kw)
        self.action_dtype = action_dtype
        self.reward_dtype = reward_dtype
        self.obs_dtype = obs_dtype
        self.record_episode_statistics = bool(record_episode_statistics)
        self.max_episode_frames = int(max_episode_frames)
        self.num_actions = 13 if simplify_action else 18
        n = self.num_envs
        self.host_mapped = bool(host_mapped)

        def zeros(shape, dtype):
            if self.host_mapped:  # cudaHostAlloc'ed by torch: device-accessible at the same address
                return torch.zeros(shape, dtype=dtype).pin_memory()
            return torch.zeros(shape, dtype=dtype, device=self.device)

        self._io_device = torch.device("cpu") if self.host_mapped else self.device
        with torch.cuda.device(self.device):
            # the packed state stays in device memory in every mode (a mapped state would put two PCIe round trips
            # in front of every frame); host_mapped maps the per-call inputs and outputs only
            self.state = torch.zeros(_lib.STATE_WORDS * n, dtype=torch.int32, device=self.device)
            self.obs_layout = obs_layout
            if obs_layout == "feature_major":
                self.obs = zeros((2, int(obs_feature_rows), n), obs_dtype)
            elif obs_layout == "shared":
                self.obs = zeros((n, _lib.OBS_WORDS), obs_dtype)
            else:
                self.obs = zeros((n, 2, _lib.OBS_WORDS), obs_dtype)
            self.reward = zeros((n, 2), reward_dtype)
            self.done_u8 = zeros((n,), torch.uint8)
            self.stats = torch.zeros(_lib.NUM_STATS, dtype=torch.int64, device=self.device) if track_stats else None
            self.episode_return = self.episode_length = self._truncated_u8 = None
            self._ep = None
            self.status = None
            if self.record_episode_statistics or self.max_episode_frames > 0 or status:
                self._ep = _lib.PzEpisodeIo()
                if status:
                    self.status = zeros((n,), torch.uint8)
                    self._ep.status_dev = self.status.data_ptr()
                if self.record_episode_statistics:
                    self.episode_return = zeros((n, 2), torch.float64)
                    self.episode_length = zeros((n,), torch.int32)
                    self._ep.episode_return_dev = self.episode_return.data_ptr()
                    self._ep.episode_length_dev = self.episode_length.data_ptr()
                if self.max_episode_frames > 0:
                    self._truncated_u8 = zeros((n,), torch.uint8)
                    self._ep.truncated_dev = self._truncated_u8.data_ptr()
            # host_mapped, at most one CTA of envs: the step kernel stores a completion word behind a system-scope
            # fence, so the host can spin on it (`wait()`) instead of synchronising the stream — half of the latency
            # of a host-driven single-env step
            self._np_seq, self._seq_count = None, 0
            if self.host_mapped and n <= 128:
                if self._ep is None:
                    self._ep = _lib.PzEpisodeIo()
                self._seq = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._np_seq = self._seq.numpy()
                self._ep.seq_dev = self._seq.data_ptr()
            _lib.check(
                self.lib.pz_seed(self.state.data_ptr(), n, self.seed & (2**64 - 1), self.first_env, self._stream()),
                "pz_seed",
            )
            # The memoised trajectory tables of the computer players (1.9 GB per device, ~0.3 s to build) are made
            # here, explicitly, so that no step() pays for them, CUDA-graph capture of step() works, and a failure
            # is reported with its cause instead of silently falling back to the iterative simulations.
            self.tables_ready = False
            if self.landing_tables and (is_player1_computer or is_player2_computer):
                rc = self.lib.pz_tables_prepare(self._stream())
                if rc != 0:
                    import warnings

                    warnings.warn(f"landing tables unavailable ({self.lib.pz_strerror(rc).decode()}, "
                                  f"{self.lib.pz_tables_bytes() / 1e9:.1f} GB needed): computer players run the "
                                  "iterative simulations (identical results, slower)", RuntimeWarning, stacklevel=2)
                self.tables_ready = rc == 0
        self._renderer = None
        self._checked_actions, self._checked_ptr = object(), None
        self.frame = 0  # calls issued so far (drives the synthetic action stream of rollout())
        self._action_shape = torch.Size((n, 2))
        self._step_args = None
        self._step_result = (self.obs, self.reward, self.done_u8.view(torch.bool))

    # ---- plumbing ------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _cfg_ref(self):
        return ctypes.byref(self.cfg)

    def _stats_ptr(self):
        return self.stats.data_ptr() if self.stats is not None else None

    def _ep_ref(self):
        return ctypes.byref(self._ep) if self._ep is not None else None

    @property
    def truncated(self) -> Optional[torch.Tensor]:
        """[N] bool (None unless max_episode_frames > 0): the episode hit the frame cap on the last call."""
        return self._truncated_u8.view(torch.bool) if self._truncated_u8 is not None else None

    # ---- rendering ---------------------------------------------------------------------------
    def attach_renderer(self, indices, sprite_dir=None, cloud_seed: int = 0, sprites=None):
        """rgb_array frames (raw_env.render(), pikazoo_env.py:250-384) for the envs `indices` of this batch:
        returns a `pikazoo_b200.render.BatchRenderer` whose `render()` gives a uint8 CUDA tensor
        [len(indices), 304, 432, 3]. While attached, every reset() / step() exports the selected envs' states to the
        host (the render-only ball state is tracked there), so attach it for evaluation runs, not for training."""
        from .render import BatchRenderer

        self._renderer = BatchRenderer(self, indices, sprite_dir=sprite_dir, cloud_seed=cloud_seed, sprites=sprites)
        return self._renderer

    def detach_renderer(self) -> None:
        self._renderer = None

    # ---- reference-shaped API --------------------------------------------------------------
    def reset(self) -> torch.Tensor:
        """reference reset() on every env; returns obs [N, 2, 35] int32 (a view reused by step)."""
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.pz_reset_ex(self.state.data_ptr(), self.num_envs, self._cfg_ref(), self.obs.data_ptr(),
                                     self._ep_ref(), self._stream()),
                "pz_reset",
            )
            if self._truncated_u8 is not None:
                self._truncated_u8.zero_()
        if self._renderer is not None:
            self._renderer.after_reset()
        return self.obs

    def step(self, actions: Optional[torch.Tensor]):
        """actions [N, 2] (int32/int64/uint8 as configured) -> (obs, reward [N,2], terminated [N] bool).

        The returned tensors are the env's own output buffers, overwritten by the next call.
        `actions` may be None only when both players are computers.
        """
        if actions is self._checked_actions:  # the same tensor object as last time (its storage cannot have moved)
            a_ptr = self._checked_ptr
        elif actions is not None:
            if actions.dtype != self.action_dtype:
                raise TypeError(f"actions must be {self.action_dtype} (got {actions.dtype}); "
                                "pass action_dtype= to the constructor")
            if actions.device != self._io_device or actions.shape != self._action_shape:
                raise ValueError(f"actions must be a [{self.num_envs}, 2] tensor on {self._io_device}")
            if not actions.is_contiguous():
                # (a .contiguous() copy here would be a temporary freed while the launch is in flight, and for
                # host_mapped a pageable one the device cannot address)
                raise ValueError("actions must be contiguous")
            if self.host_mapped and not actions.is_pinned():
                raise ValueError("host_mapped: actions must live in pinned host memory (tensor.pin_memory())")
            a_ptr = actions.data_ptr()
            if self.host_mapped:  # is_pinned() asks the driver (1-2 us): remember the verdict for this object
                self._checked_actions, self._checked_ptr = actions, a_ptr
        else:
            a_ptr = None
        if self._renderer is not None:
            self._renderer.before_call()
        if self._np_seq is not None:
            self._seq_count = (self._seq_count + 1) & 0x7FFFFFFF
            self._ep.seq_value = self._seq_count
        # the small-batch regime is bound by this host path: every constant argument is cached
        # (self._step_args), and the device guard is only entered when another device is current
        args = self._step_args
        if args is None:
            args = self._step_args = (self.state.data_ptr(), self.num_envs, self._cfg_ref(), self.obs.data_ptr(),
                                      self.reward.data_ptr(), self.done_u8.data_ptr(), self._stats_ptr(),
                                      self._ep_ref())
        if _raw_device() == self._dev_index:
            rc = self.lib.pz_step_ex(args[0], args[1], args[2], a_ptr, args[3], args[4], args[5], args[6], args[7],
                                     _raw_stream(self._dev_index))
        else:
            with torch.cuda.device(self.device):
                rc = self.lib.pz_step_ex(args[0], args[1], args[2], a_ptr, args[3], args[4], args[5], args[6],
                                         args[7], self._stream())
        if rc != 0:
            _lib.check(rc, "pz_step")
        self.frame += 1
        if self._renderer is not None:
            self._renderer.after_step()
        return self._step_result

    def wait(self) -> None:
        """Block until the last step() has written all of its outputs. host_mapped batches of at most 128 envs spin
        on the kernel's completion word; everything else synchronises the stream."""
        if self._np_seq is not None and self._seq_count:
            seq, want = self._np_seq, self._seq_count
            for _ in range(2_000_000):
                if seq[0] == want:
                    return
        torch.cuda.current_stream(self.device).synchronize()

    def rollout(self, K: int, actions: str = "noop", action_seed: int = 0, write_obs: bool = False):
        """K frames in one launch with the state in registers (auto-reset always on).

        actions: "noop" (both 0; computer players decide for themselves) or "synth" (uniform
        actions from the counter-based device stream keyed by (action_seed, global env, frame)).
        """
        if self._renderer is not None:
            raise RuntimeError("a renderer follows the env call by call: detach it before a K-frame rollout")
        src = {"noop": _lib.ACTIONS_NOOP, "synth": _lib.ACTIONS_SYNTH}[actions]
        with torch.cuda.device(self.device):
            _lib.check(
                self.lib.pz_rollout(self.state.data_ptr(), self.num_envs, self._cfg_ref(), int(K), src,
                                    int(action_seed) & (2**64 - 1), self.first_env, self.frame,
                                    self.obs.data_ptr() if write_obs else None, self._stats_ptr(), self._stream()),
                "pz_rollout",
            )
        self.frame += int(K)
        return self.obs if write_obs else None

    # ---- state access ------------------------------------------------------------------------
    def export_state(self) -> torch.Tensor:
        """Unpacked parity state int32 [N, 53] (layout: oracle/pika_oracle.h pk_env)."""
        out = torch.empty((self.num_envs, _lib.UNPACKED_WORDS), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pz_export_state(self.state.data_ptr(), self.num_envs, out.data_ptr(), self._stream()),
                       "pz_export_state")
        return out

    def import_state(self, unpacked: torch.Tensor) -> None:
        u = unpacked.to(device=self.device, dtype=torch.int32).contiguous()  # staged on the device in every mode
        assert tuple(u.shape) == (self.num_envs, _lib.UNPACKED_WORDS)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pz_import_state(self.state.data_ptr(), self.num_envs, u.data_ptr(), self._stream()),
                       "pz_import_state")

    def scores(self) -> torch.Tensor:
        """[N, 2] current scores (info["score"] of the reference, pikazoo_env.py:573-574)."""
        return self.export_state()[:, 37:39]

    def stats_dict(self) -> dict:
        """Episode statistics accumulated on the device (one D2H sync)."""
        if self.stats is None:
            return {}
        v = self.stats.tolist()
        d = {name: int(v[i]) for i, name in enumerate(_lib.STAT_NAMES)}
        d["env_steps"] = d["calls"] - d["resets"] - d["frozen"]
        return d

    def state_dict(self) -> dict:
        """The packed SoA state tensor is the checkpoint."""
        sd = {"state": self.state.clone(), "frame": self.frame, "kw": dict(self._kw), "seed": self.seed,
              "first_env": self.first_env, "num_envs": self.num_envs}
        if self.episode_return is not None:
            sd["episode_return"] = self.episode_return.clone()
        return sd

    def load_state_dict(self, sd: dict) -> None:
        assert sd["num_envs"] == self.num_envs
        self.state.copy_(sd["state"])
        self.frame = int(sd["frame"])
        if self.episode_return is not None and "episode_return" in sd:
            self.episode_return.copy_(sd["episode_return"])
