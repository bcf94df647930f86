"""RewardInNormalState (reference: pikazoo/wrappers/reward_in_normal_state.py:5-15): every agent whose
reward is 0 on a step gets `reward` instead. Fused into the kernel's reward table; the position of this
wrapper relative to RewardByBallPosition is kept (inside it if constructed first, outside otherwise)."""

from __future__ import annotations

from .base import BaseParallelWrapper


class RewardInNormalState(BaseParallelWrapper):
    def __init__(self, env, reward):
        super().__init__(env)
        self.reward = reward
        raw = env.unwrapped
        # constructed before any RewardByBallPosition => it is the inner wrapper of the two
        self._fused = raw._try_fuse("rins", reward_in_normal_state=reward,
                                    normal_state_first=raw._reward_by_ball_position is None)

    def step(self, actions):
        res = self.env.step(actions)
        if not self._fused:  # over RecordEpisodeStatistics or another RewardInNormalState: on the host
            rews = res[1]
            for agent in self.possible_agents:
                if rews[agent] == 0:
                    rews[agent] = self.reward
        return res
