"""The two wrappers north_star names, with the reference's constructor signatures
(pikazoo/wrappers/simplify_action.py:7-28, reward_by_ball_position.py:6-31). They do not
post-process in Python: they switch on the fused path of the step kernel."""

from .simplify_action import SimplifyAction
from .reward_by_ball_position import RewardByBallPosition

__all__ = ["SimplifyAction", "RewardByBallPosition"]
