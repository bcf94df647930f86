"""A/B timer of the kernels that are bound by instruction issue rather than by HBM, in steady state: one JSON line per
library build (PIKAZOO_B200_LIB selects it; python pikazoo_b200/build.py --out variants/NAME.so builds one).
    python profiles/ab_rollout.py [reps]
  rollout64_ai_us        configs[3]: 1,048,576 envs computer vs computer, K = 64 frames per launch (pz_rollout)
  rollout64_synth_us     the same launch without computer players, synthetic actions
  step_ai_us             computer vs computer through the per-step kernel
  rollout_policy64_us    configs[4]: 2,097,152 envs, MLP policy inside the K = 64 launch (pz_rollout_policy)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import MLPPolicy, rollout_fused  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << 20
out = {"lib": os.environ.get("PIKAZOO_B200_LIB", "product"), "reps": reps}


def best_of(fn, reps=reps, rounds=3):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e18
    for _ in range(rounds):
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1) * 1e3 / reps)
    return round(best, 2)


kw = dict(winning_score=15, serve="winner")
e = pikazoo_b200.PikaVecEnv(n, seed=3, is_player1_computer=True, is_player2_computer=True, **kw)
e.reset()
for _ in range(8):
    e.rollout(256)
out["rollout64_ai_us"] = best_of(lambda: e.rollout(64))
out["rollout64_ai_env_steps_per_s"] = round(n * 64 / (out["rollout64_ai_us"] * 1e-6) / 1e9, 2)
out["step_ai_us"] = best_of(lambda: e.step(None), reps * 20)
del e
e = pikazoo_b200.PikaVecEnv(n, seed=5, **kw)
e.reset()
for _ in range(8):
    e.rollout(256, actions="synth", action_seed=77)
out["rollout64_synth_us"] = best_of(lambda: e.rollout(64, actions="synth", action_seed=7))
del e
env = pikazoo_b200.PikaVecEnv(2 * n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                              normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                              obs_feature_rows=40)
pol = MLPPolicy(device=env.device, seed=3)
env.reset()
for _ in range(4):
    rollout_fused(env, pol, 64, seed=1)
out["rollout_policy64_us"] = best_of(lambda: rollout_fused(env, pol, 64, seed=1), max(reps // 3, 5))
out["rollout_policy_env_steps_per_s"] = round(2 * n * 64 / (out["rollout_policy64_us"] * 1e-6) / 1e9, 2)
print(json.dumps(out))
