"""ncu driver (use with --profile-from-start off): ONE steady-state launch each of
   0: pz_step_kernel<0, F16, ENV_MAJOR>      1,048,576 envs, normalised rows, uint8 actions (bench variant)
   1: pz_step_kernel<0, BF16, FEATURE_MAJOR> 2,097,152 envs, configs[4]'s env step (ws 5, serve random)
   2: pz_step_kernel<0, I32, ENV_MAJOR>      1,048,576 envs, the bench's main kernel
   3: pz_rollout_policy_kernel<18>           524,288 envs, K = 16
   4: pz_rollout_kernel<3>                   524,288 envs, K = 16, computer vs computer
between cudaProfilerStart/Stop; everything else (reset, 2,048 frames of pre-advance, warm-up) is outside.
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/hot python profiles/profile_hotspots.py
    python profiles/source_hotspots.py gpurun_out/hot.ncu-rep <index> 50"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import MLPPolicy, rollout_fused  # noqa: E402

which = set(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else {0, 1, 2, 3, 4}
prof = torch.cuda.profiler
g = torch.Generator(device="cuda").manual_seed(1)


def steady(env, ring):
    env.reset()
    for _ in range(8):
        env.rollout(256, actions="synth", action_seed=77)
    for k in range(6):
        env.step(ring[k % len(ring)])
    torch.cuda.synchronize()
    prof.start()
    env.step(ring[0])
    torch.cuda.synchronize()
    prof.stop()


n = 1 << 20
if 0 in which:
    ring = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.uint8) for _ in range(4)]
    steady(pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.float16, normalize_observation=True, action_dtype=torch.uint8,
                                   winning_score=15, serve="winner"), ring)
if 1 in which:
    ring = [torch.randint(0, 18, (2 * n, 2), generator=g, device="cuda", dtype=torch.uint8) for _ in range(4)]
    steady(pikazoo_b200.PikaVecEnv(2 * n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                                   normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                                   obs_feature_rows=40), ring)
if 2 in which:
    ring = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(4)]
    steady(pikazoo_b200.PikaVecEnv(n, seed=2026, winning_score=15, serve="winner"), ring)
if 3 in which:
    env = pikazoo_b200.PikaVecEnv(n // 2, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                                  normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major",
                                  obs_feature_rows=40)
    pol = MLPPolicy(device=env.device, seed=3)
    env.reset()
    for _ in range(4):
        rollout_fused(env, pol, 64, seed=1)
    torch.cuda.synchronize()
    prof.start()
    rollout_fused(env, pol, 16, seed=1)
    torch.cuda.synchronize()
    prof.stop()
if 4 in which:
    env = pikazoo_b200.PikaVecEnv(n // 2, seed=3, is_player1_computer=True, is_player2_computer=True, winning_score=15,
                                  serve="winner")
    env.reset()
    for _ in range(8):
        env.rollout(64)
    torch.cuda.synchronize()
    prof.start()
    env.rollout(16)
    torch.cuda.synchronize()
    prof.stop()
print("ok")
