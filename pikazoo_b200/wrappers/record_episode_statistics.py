"""RecordEpisodeStatistics (reference: pikazoo/wrappers/record_episode_statistics.py:9-40): on the step
that ends an episode, infos[agent]["episode"] = {"r": episode return, "l": episode length}. Fused: the
kernel keeps the per-env running returns (float64, added in step order) and lengths."""

from __future__ import annotations

from .base import BaseParallelWrapper


class RecordEpisodeStatistics(BaseParallelWrapper):
    def __init__(self, env):
        super().__init__(env)
        env.unwrapped._configure(record_episode_statistics=True)

    @property
    def episode_rewards(self):
        return self.env.unwrapped._episode_rewards()

    @property
    def episode_lengths(self):
        return self.env.unwrapped._episode_lengths()
