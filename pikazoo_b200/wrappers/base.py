"""Delegating wrapper base (stand-in for pettingzoo.utils.BaseParallelWrapper)."""

from __future__ import annotations


class BaseParallelWrapper:
    def __init__(self, env):
        self.env = env
        self.metadata = getattr(env, "metadata", {})
        self.possible_agents = env.possible_agents

    @property
    def agents(self):
        return self.env.agents

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def scores(self):
        return self.env.scores

    def reset(self, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, actions):
        return self.env.step(actions)

    def observation_space(self, agent):
        return self.env.observation_space(agent)

    def action_space(self, agent):
        return self.env.action_space(agent)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()
