"""RewardByBallPosition (reference: pikazoo/wrappers/reward_by_ball_position.py:6-31): adds
additional_reward[4*agent + zone] to each agent's reward, zone = (ball_y > y_line) + 2*(ball_x >= x_line)
from the post-step ball. Fused: the kernel reads an [agent][base reward][zone] table built on the
host in double precision, so results equal Python's `int + float`."""

from __future__ import annotations

from .base import BaseParallelWrapper


class RewardByBallPosition(BaseParallelWrapper):
    def __init__(self, env, additional_reward, x_line: int = 216, y_line: int = 176):
        super().__init__(env)
        assert len(additional_reward) == 8
        self.x_line = x_line
        self.y_line = y_line
        self.additional_reward = additional_reward
        self._fused = env.unwrapped._try_fuse(
            "rbbp", reward_by_ball_position=(tuple(additional_reward), int(x_line), int(y_line)))

    def step(self, actions):
        res = self.env.step(actions)
        if self._fused:
            return res
        # not fusable at this position (over NormalizeObservation, RecordEpisodeStatistics or another
        # RewardByBallPosition): reward_by_ball_position.py:20-31 on the host, on the observation of the env below
        obs, rews = res[0], res[1]
        ball_x, ball_y = obs["player_1"][26], obs["player_1"][27]
        zone = int(ball_y > self.y_line) + 2 * int(ball_x >= self.x_line)
        for i, agent in enumerate(self.possible_agents):
            rews[agent] += self.additional_reward[i * 4 + zone]
        return res
