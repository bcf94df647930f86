"""pikazoo_b200: B200-native batched Pikachu Volleyball (drop-in for the hot path of
helpingstar/pika-zoo).

    import pikazoo_b200
    from pikazoo_b200 import pikazoo_v0, PikaVecEnv
    from pikazoo_b200.wrappers import SimplifyAction
"""

from . import pikazoo_v0, wrappers  # noqa: F401
from ._lib import PikaLibraryError, load as load_library  # noqa: F401
from .dist import allreduce_stats, make_sharded_env, shard_range  # noqa: F401
from .vec_env import PikaVecEnv, make_config  # noqa: F401
from .wrappers import (ConvertSingleAgent, NormalizeObservation, RecordEpisodeStatistics,  # noqa: F401
                       RewardByBallPosition, RewardInNormalState, SimplifyAction)

__version__ = "0.1.0"
