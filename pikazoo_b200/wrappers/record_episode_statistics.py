"""RecordEpisodeStatistics (reference: pikazoo/wrappers/record_episode_statistics.py:9-40): on the step
that ends an episode, infos[agent]["episode"] = {"r": episode return, "l": episode length}. Fused: the
kernel keeps the per-env running returns (float64, added in step order) and lengths."""

from __future__ import annotations

from .base import BaseParallelWrapper


class RecordEpisodeStatistics(BaseParallelWrapper):
    def __init__(self, env):
        super().__init__(env)
        self._fused = env.unwrapped._try_fuse("record", record_episode_statistics=True)
        self._rewards = {agent: 0 for agent in self.possible_agents}
        self._lengths = {agent: 0 for agent in self.possible_agents}

    @property
    def episode_rewards(self):
        return self.env.unwrapped._episode_rewards() if self._fused else self._rewards

    @property
    def episode_lengths(self):
        return self.env.unwrapped._episode_lengths() if self._fused else self._lengths

    def reset(self, seed=None, options=None):
        res = self.env.reset(seed=seed, options=options)
        for agent in self.possible_agents:
            self._rewards[agent] = 0
            self._lengths[agent] = 0
        return res

    def step(self, actions):
        res = self.env.step(actions)
        if self._fused:
            return res
        # a second RecordEpisodeStatistics: record_episode_statistics.py:27-40 on the host
        obs, rews, terminateds, truncateds, infos = res
        for agent in self.possible_agents:
            self._rewards[agent] += rews[agent]
            self._lengths[agent] += 1
        if all(terminateds.values()) or all(truncateds.values()):
            for agent in self.possible_agents:
                infos[agent] = dict(infos[agent], episode={"r": self._rewards[agent], "l": self._lengths[agent]})
        return res
