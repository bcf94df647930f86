"""configs[4] (2,097,152 envs per GPU, winning_score 5, serve random, both agents' actions from the MLP policy) through
pz_rollout_policy — K frames per launch, everything on chip — next to the two-kernel loop (pz_policy_mlp_act + pz_step).

    python profiles/time_rollout_policy.py [--envs N] [--K 64] [--reps 5] [--no-loop] [--export-actions]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout, rollout_fused  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 21)
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-loop", action="store_true")
    ap.add_argument("--export-actions", action="store_true")
    a = ap.parse_args()
    n, K = a.envs, a.K
    kw = dict(seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True,
              action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40)
    env = pikazoo_b200.PikaVecEnv(n, **kw)
    pol = MLPPolicy(device=env.device, seed=3)
    env.reset()
    acts = torch.empty((K, n, 2), dtype=torch.uint8, device="cuda") if a.export_actions else None
    for _ in range(2):
        rollout_fused(env, pol, K, seed=1, actions_out=acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        rollout_fused(env, pol, K, seed=1, actions_out=acts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    out = {"envs": n, "K": K, "export_actions": bool(a.export_actions), "fused_ms_per_launch": ms,
           "fused_us_per_frame": ms * 1e3 / K, "fused_env_steps_per_sec": n * K / (ms * 1e-3), "stats": env.stats_dict()}
    if not a.no_loop:
        env2 = pikazoo_b200.PikaVecEnv(n, **kw)
        env2.reset()
        actor = FusedActor(pol, env2, seed=1)
        policy_rollout(env2, actor, 20)
        torch.cuda.synchronize()
        e0.record()
        policy_rollout(env2, actor, 100)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 100
        out.update(two_kernel_loop_ms_per_step=ms2, two_kernel_loop_env_steps_per_sec=n / (ms2 * 1e-3),
                   speedup=ms2 / (ms / K))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
