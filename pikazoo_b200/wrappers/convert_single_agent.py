"""ConvertSingleAgent (reference: pikazoo/wrappers/convert_single_agent.py:5-28): a single-agent view of
one side; the other side plays uniformly random actions sampled from its action space. Pure host-side
plumbing, as upstream (the sampled action enters the kernel like any other)."""

from __future__ import annotations

from .base import BaseParallelWrapper


class ConvertSingleAgent(BaseParallelWrapper):
    def __init__(self, env, side: str):
        super().__init__(env)
        assert side in ("player_1", "player_2")  # convert_single_agent.py:8
        self.side = side
        self.other_side = "player_1" if side == "player_2" else "player_2"

    def reset(self, seed=None, options=None):
        obs, infos = super().reset(seed=seed, options=options)
        return obs[self.side], infos[self.side]

    def step(self, action):
        actions = {self.side: action, self.other_side: self.action_space(self.other_side).sample()}
        obs, rews, terminateds, truncateds, infos = super().step(actions)
        return obs[self.side], rews[self.side], terminateds[self.side], truncateds[self.side], infos[self.side]
