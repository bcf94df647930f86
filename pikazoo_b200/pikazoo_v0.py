"""`pikazoo_v0.env / raw_env / parallel_env`: the reference's single-env PettingZoo-style
surface (pikazoo/pikazoo_v0.py:1-3, pikazoo/env/pikazoo_env.py:27-29,72-240,481-574) on top of
a 1-env batch of the CUDA simulator. Same kwargs, agents, spaces, dict-in / five-dicts-out
protocol and termination bookkeeping; observations come back as numpy arrays like the
reference's. `parallel_env` does not exist upstream and is an alias of `env`.

Differences, all deliberate: `reset(seed=...)` DOES seed the env (the reference ignores it,
pikazoo_env.py:149); render_mode is None or "rgb_array" (no display; the sprites are the reference's
assets: pass sprite_dir= or set PIKAZOO_SPRITE_DIR; clouds and waves animate on their own generator,
so rendering does not change the game as it does upstream); an out-of-range action raises IndexError
(the reference lets numpy wrap negative indices).
"""

from __future__ import annotations

import functools
from typing import Dict, List, Optional

import numpy as np
import torch

from . import spaces
from .vec_env import AGENTS, PikaVecEnv

__all__ = ["env", "raw_env", "parallel_env"]

# observation_space bounds, pikazoo_env.py:481-565
_P_LOW = [32, 108, -15, -1, -2, 0, 0, 0, 0, 0, 0, 0, 0]
_P_HIGH = [400, 244, 16, 1, 3, 4, 4, 1, 1, 1, 1, 1, 1]
_B_LOW = [20, 0, 0, 0, 0, 0, -20, -124, 0]
_B_HIGH = [432, 252, 432, 252, 432, 252, 20, 124, 1]
OBS_LOW = np.array(_P_LOW + _P_LOW + _B_LOW)
OBS_HIGH = np.array(_P_HIGH + _P_HIGH + _B_HIGH)


def env(**kwargs):
    return raw_env(**kwargs)


def parallel_env(**kwargs):
    return raw_env(**kwargs)


class raw_env:
    metadata = {"render_modes": ["human", "rgb_array"], "name": "pikazoo_v0", "render_fps": 20}

    def __init__(
        self,
        winning_score=15,
        serve="winner",
        is_player1_computer=False,
        is_player2_computer=False,
        render_mode=None,
        device="cuda",
        seed: Optional[int] = None,
        sprite_dir: Optional[str] = None,
        cloud_seed: int = 0,
    ):
        assert serve in ("winner", "alternate", "random")  # pikazoo_env.py:104
        if render_mode not in (None, "rgb_array"):
            raise NotImplementedError("render_mode must be None or 'rgb_array' (there is no display here)")
        self.possible_agents: List[str] = list(AGENTS)
        self.agents: List[str] = self.possible_agents[:]
        self.action_spaces = dict(zip(self.agents, [spaces.Discrete(18)] * 2))
        self.winning_score = winning_score
        self.serve = serve
        self.render_mode = render_mode
        self._sprite_dir, self._cloud_seed, self._sprites = sprite_dir, int(cloud_seed), None
        if render_mode is not None:  # the reference loads its images in the constructor too (pikazoo_env.py:146-147)
            from .render import SpriteSet

            self._sprites = SpriteSet(sprite_dir)
        self._device = device
        self._kwargs = dict(
            winning_score=winning_score, serve=serve, is_player1_computer=is_player1_computer,
            is_player2_computer=is_player2_computer,
        )
        # fused wrapper options, set by pikazoo.wrappers.* before the first reset
        self._simplify_action = False
        self._reward_by_ball_position = None
        self._reward_in_normal_state = None
        self._normal_state_first = False
        self._normalize_observation = False
        self._record_episode_statistics = False
        self._stack: List[tuple] = []  # (kind, fused) of every wrapper constructed over this env, innermost first
        self._seed_value = int(np.random.SeedSequence().entropy % (2**63)) if seed is None else int(seed)
        self._vec: Optional[PikaVecEnv] = None
        self.scores: List[int] = [0, 0]
        self._actions = None

    # -- internals ---------------------------------------------------------------------------
    def _build(self):
        # host_mapped: state, observations, rewards, ... of the one env live in pinned host memory that the
        # kernel addresses directly, so a step is one launch and ONE stream synchronisation, no copies
        self._vec = PikaVecEnv(
            1, device=self._device, seed=self._seed_value, autoreset=False, reward_dtype=torch.float64,
            simplify_action=self._simplify_action, reward_by_ball_position=self._reward_by_ball_position,
            reward_in_normal_state=self._reward_in_normal_state, normal_state_first=self._normal_state_first,
            normalize_observation=self._normalize_observation,
            obs_dtype=torch.float64 if self._normalize_observation else torch.int32,
            record_episode_statistics=self._record_episode_statistics, track_stats=False, host_mapped=True,
            status=True,
            **self._kwargs,
        )
        v = self._vec
        if self.render_mode is not None:
            v.attach_renderer([0], cloud_seed=self._cloud_seed, sprites=self._sprites)
        self._actions = torch.zeros((1, 2), dtype=torch.int32).pin_memory()
        # numpy views of the mapped buffers, read after the synchronisation
        self._np_actions = self._actions.numpy()
        self._np_obs = v.obs.numpy()
        self._np_reward = v.reward.numpy()
        self._np_done = v.done_u8.numpy()
        self._np_status = v.status.numpy()
        self._sync = torch.cuda.current_stream(v.device).synchronize

    # What the kernel can reproduce depends on WHERE a wrapper sits in the stack (wrappers are constructed inside
    # out): RecordEpisodeStatistics records the rewards of what is below it, RewardByBallPosition reads the
    # observation of what is below it (normalised values if NormalizeObservation is), two RewardByBallPosition add
    # up, ... The fused options implement ONE order — SimplifyAction, RewardInNormalState / RewardByBallPosition in
    # either order, NormalizeObservation, RecordEpisodeStatistics on top of the reward wrappers. A wrapper that
    # does not fit does its work on the host, exactly as the reference's class does, and so does everything
    # stacked on top of it.
    _FUSABLE_OVER = {
        "simplify": lambda below: "simplify" not in below,
        "rbbp": lambda below: not ({"rbbp", "normalize", "record"} & below),
        "rins": lambda below: not ({"rins", "record"} & below),
        "normalize": lambda below: "normalize" not in below,
        "record": lambda below: "record" not in below,
    }

    def _try_fuse(self, kind: str, **opts) -> bool:
        """Called by a wrapper's constructor. True: the kernel option is switched on and the wrapper is a pass-through.
        False: the wrapper must post-process on the host."""
        fused = all(f for _, f in self._stack) and self._FUSABLE_OVER[kind]({k for k, _ in self._stack})
        self._stack.append((kind, fused))
        if fused:
            self._configure(**opts)
        return fused

    def _configure(self, **opts):
        """Used by the wrappers to fuse themselves into the kernel configuration."""
        for k, v in opts.items():
            setattr(self, "_" + k, v)
        if self._vec is not None:  # keep the live state, swap the config
            state = self._vec.state_dict()
            self._build()
            self._vec.load_state_dict(state)

    def _fetch(self, stepped=False):
        """wait for the launch; the outputs are then in the mapped buffers"""
        if stepped:
            self._vec.wait()  # spins on the kernel's completion word in mapped memory
        else:
            self._sync()
        # the scores follow from the status byte: its low two bits are player_1's BASE reward + 1 (whatever reward
        # wrappers are fused), and a point is scored exactly when that is non-zero (pikazoo_env.py:190-223)
        if stepped:
            base = (int(self._np_status[0]) & 3) - 1
            if base > 0:
                self.scores[0] += 1
            elif base < 0:
                self.scores[1] += 1
        else:
            self.scores[0] = self.scores[1] = 0

    def _obs_dict(self) -> Dict[str, np.ndarray]:
        # the reference returns np.array of Python ints (int64); NormalizeObservation makes them float64
        o = self._np_obs[0]
        o = o.copy() if self._normalize_observation else o.astype(np.int64)
        return {self.possible_agents[0]: o[0], self.possible_agents[1]: o[1]}

    def _get_infos(self):
        return {agent: {"score": self.scores} for agent in self.agents}

    # -- reference API ---------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            self._seed_value = int(seed)
            self._vec = None
        if self._vec is None:
            self._build()
        self.agents = self.possible_agents[:]
        self._vec.reset()
        self._fetch()
        return self._obs_dict(), self._get_infos()

    def step(self, actions):
        if not self.agents:
            raise IndexError("step() called on a terminated env; call reset() (reference: pikazoo_env.py:237-238)")
        if self._vec is None:
            raise RuntimeError("call reset() before step()")
        n = 13 if self._simplify_action else 18
        a = [int(actions[agent]) for agent in self.agents]
        for v in a:
            if not 0 <= v < n:
                raise IndexError(f"action {v} is out of range for Discrete({n})")
        self._np_actions[0, 0], self._np_actions[0, 1] = a[0], a[1]
        self._vec.step(self._actions)
        self._fetch(stepped=True)
        r = self._np_reward[0].tolist()
        terminated = bool(self._np_done[0])
        observations = self._obs_dict()
        if self._reward_by_ball_position is None and not isinstance(self._reward_in_normal_state, float):
            r = [int(r[0]), int(r[1])]  # the reference's base rewards are Python ints
        rewards = {self.agents[0]: r[0], self.agents[1]: r[1]}
        terminations = {agent: terminated for agent in self.agents}
        truncations = {agent: False for agent in self.agents}
        infos = self._get_infos()
        if terminated and self._record_episode_statistics:  # record_episode_statistics.py:34-39
            ret, length = self._episode_rewards(), self._episode_lengths()
            for agent in self.agents:
                infos[agent] = dict(infos[agent], episode={"r": ret[agent], "l": length[agent]})
        if terminated:
            self.agents = []
        return observations, rewards, terminations, truncations, infos

    def _episode_rewards(self):
        r = self._vec.episode_return[0].cpu().tolist()
        if self._reward_by_ball_position is None and not isinstance(self._reward_in_normal_state, float):
            r = [int(r[0]), int(r[1])]
        return dict(zip(self.possible_agents, r))

    def _episode_lengths(self):
        n = int(self._vec.episode_length[0].item())
        return {agent: n for agent in self.possible_agents}

    @functools.lru_cache(maxsize=None)
    def observation_space(self, agent=None):
        return spaces.Box(low=OBS_LOW, high=OBS_HIGH, shape=(35,), dtype=np.int32)

    def action_space(self, agent):
        return self.action_spaces[agent]

    def render(self):
        """uint8 [304, 432, 3] like the reference's render_mode="rgb_array" (pikazoo_env.py:368-384); None without
        a render mode. Clouds and waves animate on their own generator (cloud_seed): rendering never touches the
        game's random stream (in the reference it does)."""
        if self.render_mode is None:
            return None
        if self._vec is None:
            raise RuntimeError("call reset() before render()")
        return self._vec._renderer.render()[0].cpu().numpy()

    def close(self):
        self._vec = None

    @property
    def unwrapped(self):
        return self

    def state_words(self) -> np.ndarray:
        """52-word hidden state (oracle/pika_oracle.h layout) of this env, for tests."""
        return self._vec.export_state()[0, :52].cpu().numpy()
