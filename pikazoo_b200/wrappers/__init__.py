"""The reference's wrappers with their constructor signatures (pikazoo/wrappers/*.py). Except for
ConvertSingleAgent (host-side plumbing upstream too) they do not post-process in Python: they switch
on the corresponding fused path of the step kernel."""

from .convert_single_agent import ConvertSingleAgent
from .normalize_observation import NormalizeObservation
from .record_episode_statistics import RecordEpisodeStatistics
from .reward_by_ball_position import RewardByBallPosition
from .reward_in_normal_state import RewardInNormalState
from .simplify_action import SimplifyAction

__all__ = ["SimplifyAction", "RewardByBallPosition", "RewardInNormalState", "NormalizeObservation",
           "RecordEpisodeStatistics", "ConvertSingleAgent"]
