"""configs[4] on one GPU: ms per step of the obs -> policy -> step loop at 2 M envs, eager PyTorch policy against
the fused policy kernel (CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
kw = dict(seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True,
          obs_layout="feature_major", obs_feature_rows=40)
pol = MLPPolicy()
out = {"n_envs": n}


def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn(reps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


env = pikazoo_b200.PikaVecEnv(n, action_dtype=torch.int64, **kw)
env.reset()
policy_rollout(env, pol.act, 10)
out["eager_policy_loop_ms_per_step"] = timed(lambda r: policy_rollout(env, pol.act, r), 50)
del env
env = pikazoo_b200.PikaVecEnv(n, action_dtype=torch.uint8, **kw)
env.reset()
actor = FusedActor(pol, env, seed=1)
policy_rollout(env, actor, 10)
out["fused_policy_loop_ms_per_step"] = timed(lambda r: policy_rollout(env, actor, r), 300)
out["fused_policy_kernel_ms"] = timed(lambda r: [actor(env.obs) for _ in range(r)], 300)
acts = actor(env.obs)
out["env_step_ms"] = timed(lambda r: [env.step(acts) for _ in range(r)], 300)
out["fused_loop_env_steps_per_s"] = n / (out["fused_policy_loop_ms_per_step"] * 1e-3)
out["eager_loop_env_steps_per_s"] = n / (out["eager_policy_loop_ms_per_step"] * 1e-3)
print(json.dumps(out))
