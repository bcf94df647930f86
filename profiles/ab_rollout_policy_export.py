import json, os, sys
sys.path.insert(0, "/root/repo")
import torch, pikazoo_b200
from pikazoo_b200.policy import MLPPolicy, rollout_fused
n = 1 << 21
env = pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40)
pol = MLPPolicy(device=env.device, seed=3)
env.reset()
acts = torch.empty((64, n, 2), dtype=torch.uint8, device="cuda")
out = {}
for name, a in (("no_export", None), ("export", acts)):
    for _ in range(4): rollout_fused(env, pol, 64, seed=1, actions_out=a)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e18
    for _ in range(3):
        t0.record()
        for _ in range(8): rollout_fused(env, pol, 64, seed=1, actions_out=a)
        t1.record(); torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1) * 1e3 / 8)
    out[name + "_us"] = round(best, 1); out[name + "_G"] = round(n * 64 / (best * 1e-6) / 1e9, 2)
print(json.dumps(out))
