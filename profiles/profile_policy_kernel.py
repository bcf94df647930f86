"""ncu driver: a few launches of the fused policy kernel on 2 M envs of played feature-major bf16 observations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
env = pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                              normalize_observation=True, obs_layout="feature_major", obs_feature_rows=40,
                              action_dtype=torch.uint8)
env.reset()
actor = FusedActor(MLPPolicy(), env, seed=1)
policy_rollout(env, actor, 6)
torch.cuda.synchronize()
print("ok")
