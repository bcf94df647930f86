"""Quick A/B timer (CUDA events, after warm-up) of the kernel variants at 1,048,576 envs; one JSON line.

    python profiles/time_kernels.py [n_envs] [reps]

Not the bench: no host path, no CPU baseline. Used to compare builds (PZ_NVCC_FLAGS=... python pikazoo_b200/build.py).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
kw = dict(winning_score=15, serve="winner")
g = torch.Generator(device="cuda").manual_seed(1)
a32 = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(8)]
a8 = [a.to(torch.uint8) for a in a32]
out = {"n_envs": n, "reps": reps}


def timed(fn, reps, warm):
    for k in range(warm):
        fn(k)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(reps):
        fn(k)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) * 1e3 / reps  # us per call


e = pikazoo_b200.PikaVecEnv(n, seed=1, **kw)
e.reset()
out["step_i32_us"] = timed(lambda k: e.step(a32[k % 8]), reps, 50)
del e
e = pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.float16, normalize_observation=True, action_dtype=torch.uint8, **kw)
e.reset()
out["step_f16norm_u8_us"] = timed(lambda k: e.step(a8[k % 8]), reps, 50)
del e
e = pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.uint8,
                            obs_layout="feature_major", **kw)
e.reset()
out["step_bf16norm_u8_feature_major_us"] = timed(lambda k: e.step(a8[k % 8]), reps, 50)
del e
e = pikazoo_b200.PikaVecEnv(n, seed=3, is_player1_computer=True, is_player2_computer=True, **kw)
e.reset()
e.rollout(64)
out["step_ai_vs_ai_us"] = timed(lambda k: e.step(None), reps, 50)
out["rollout64_ai_vs_ai_us"] = timed(lambda k: e.rollout(64), max(reps // 10, 10), 5)
out["rollout64_env_steps_per_s"] = n * 64 / (out["rollout64_ai_vs_ai_us"] * 1e-6)
del e
e = pikazoo_b200.PikaVecEnv(n, seed=5, **kw)
e.reset()
out["rollout64_synth_actions_us"] = timed(lambda k: e.rollout(64, actions="synth", action_seed=7), max(reps // 10, 10), 5)
print(json.dumps(out))
