"""SimplifyAction (reference: pikazoo/wrappers/simplify_action.py:7-28): 13 relative actions
instead of 18 absolute ones, mirrored for player_2. The 2x13 remap is a LUT inside the step
kernel (csrc/pz_physics.cuh kSimplify)."""

from __future__ import annotations

from .. import spaces
from .base import BaseParallelWrapper

ACTION_MAP = {
    "player_1": (0, 1, 2, 3, 4, 6, 7, 10, 11, 12, 13, 14, 16),
    "player_2": (0, 1, 2, 4, 3, 7, 6, 10, 12, 11, 13, 15, 17),
}


class SimplifyAction(BaseParallelWrapper):
    def __init__(self, env):
        super().__init__(env)
        self.action_map = dict(ACTION_MAP)
        self.action_spaces = dict(zip(self.possible_agents, [spaces.Discrete(13)] * 2))
        self._fused = env.unwrapped._try_fuse("simplify", simplify_action=True)

    def step(self, actions):
        if not self._fused:  # a second SimplifyAction over the first, ...: map on the host like the reference does
            actions = {agent: self.action_map[agent][actions[agent]] for agent in self.possible_agents}
        return self.env.step(actions)

    def action_space(self, agent):
        return self.action_spaces[agent]
