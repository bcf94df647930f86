"""A/B timer of pz_rollout_policy on configs[4] (2,097,152 envs, K = 64), one JSON line per library build
(PIKAZOO_B200_LIB): python profiles/ab_rollout_policy.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import MLPPolicy, rollout_fused  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = 1 << 21
env = pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                              normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                              obs_feature_rows=40)
pol = MLPPolicy(device=env.device, seed=3)
env.reset()
for _ in range(6):
    rollout_fused(env, pol, 64, seed=1)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e18
for _ in range(3):
    t0.record()
    for _ in range(reps):
        rollout_fused(env, pol, 64, seed=1)
    t1.record()
    torch.cuda.synchronize()
    best = min(best, t0.elapsed_time(t1) * 1e3 / reps)
print(json.dumps({"lib": os.environ.get("PIKAZOO_B200_LIB", "product"), "rollout_policy64_us": round(best, 1),
                  "G_env_steps_per_s": round(n * 64 / (best * 1e-6) / 1e9, 2)}))
