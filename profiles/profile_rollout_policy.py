"""Small driver for ncu: a few K-frame launches of pz_rollout_policy on configs[4] (2 M envs, ws 5, serve random)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import MLPPolicy, rollout_fused  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
env = pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                              normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                              obs_feature_rows=40)
pol = MLPPolicy(device=env.device, seed=3)
env.reset()
for _ in range(launches):
    rollout_fused(env, pol, K, seed=1)
torch.cuda.synchronize()
print("ok", env.stats_dict())
