"""configs[4] loop with the fused actor, a few dozen steps — the command behind the ncu launch list
profiles/r01_launches_policy_loop.csv:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file OUT.csv \
        python profiles/launches_policy_loop.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout  # noqa: E402

n = 1 << 21
env = pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                              normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                              obs_feature_rows=40)
env.reset()
policy_rollout(env, FusedActor(MLPPolicy(), env, seed=1), 24)
torch.cuda.synchronize()
print("ok")
