"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, at sizes with
full and ragged warps / CTAs / tiles.

    compute-sanitizer --tool memcheck python profiles/sanitize_target.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
for n in (1, 100, 1024, 4096 + 8):
    for layout, rows in (("env_major", 35), ("feature_major", 40)):
        for dt, norm in ((torch.int32, False), (torch.bfloat16, True), (torch.float32, True), (torch.float64, True)):
            for ai in (False, True):
                env = pikazoo_b200.PikaVecEnv(n, seed=3, winning_score=2, serve="random", is_player1_computer=ai,
                                              is_player2_computer=ai, obs_dtype=dt, normalize_observation=norm,
                                              obs_layout=layout, obs_feature_rows=rows, landing_tables=False,
                                              record_episode_statistics=True, max_episode_frames=50)
                env.reset()
                for _ in range(6):
                    env.step(torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32))
                env.rollout(8, actions="synth", action_seed=1, write_obs=True)
                env.export_state()
    env = pikazoo_b200.PikaVecEnv(n, seed=5, obs_dtype=torch.bfloat16, normalize_observation=True,
                                  obs_layout="feature_major", obs_feature_rows=40, action_dtype=torch.uint8)
    env.reset()
    logits = torch.empty((n, 2, 18), device="cuda")
    pol = MLPPolicy()
    pol.act_fused(env.obs, step=0, logits_out=logits)
    policy_rollout(env, FusedActor(pol, env), 4)
    host = pikazoo_b200.PikaVecEnv(n, seed=5, host_mapped=True, is_player2_computer=True, landing_tables=False)
    host.reset()
    a = torch.zeros((n, 2), dtype=torch.int32).pin_memory()
    for _ in range(4):
        host.step(a)
    torch.cuda.synchronize()
torch.cuda.synchronize()
print("sanitize target done")
