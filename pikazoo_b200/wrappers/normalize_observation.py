"""NormalizeObservation (reference: pikazoo/wrappers/normalize_observation.py:8-35): observations become
(obs - low) / (high - low) with the bounds of raw_env.observation_space (pikazoo_env.py:485-562), as
float64 arrays like the reference's numpy true division. Fused: the kernel emits the normalised rows
itself (csrc/pz_physics.cuh obs_float, one correctly rounded division per element)."""

from __future__ import annotations

import numpy as np

from .. import spaces
from .base import BaseParallelWrapper


class NormalizeObservation(BaseParallelWrapper):
    def __init__(self, env):
        super().__init__(env)
        self.high = {agent: env.observation_space(agent).high for agent in self.possible_agents}
        self.low = {agent: env.observation_space(agent).low for agent in self.possible_agents}
        self._fused = env.unwrapped._try_fuse("normalize", normalize_observation=True)

    def _normalize(self, obs):
        for agent in self.possible_agents:
            obs[agent] = (obs[agent] - self.low[agent]) / (self.high[agent] - self.low[agent])

    def reset(self, seed=None, options=None):
        obs, infos = self.env.reset(seed=seed, options=options)
        if not self._fused:  # a second NormalizeObservation, ...: normalize_observation.py:18-32 on the host
            self._normalize(obs)
        return obs, infos

    def step(self, actions):
        res = self.env.step(actions)
        if not self._fused:
            self._normalize(res[0])
        return res

    def observation_space(self, agent):
        return spaces.Box(low=0.0, high=1.0, shape=(35,), dtype=np.float32)  # normalize_observation.py:34-35
