"""A/B timer of the per-step kernels in STEADY STATE (every env advanced 2,048 frames, as bench.py does): one JSON line
per library build. PIKAZOO_B200_LIB selects the build (python pikazoo_b200/build.py --out variants/NAME.so).
    python profiles/ab_step_steady.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n = 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
a32 = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(8)]
a8 = [a.to(torch.uint8) for a in a32]
out = {"lib": os.environ.get("PIKAZOO_B200_LIB", "product"), "reps": reps}


def steady(env, ring, reps=reps):
    env.reset()
    for _ in range(8):
        env.rollout(256, actions="synth", action_seed=77)
    for k in range(100):
        env.step(ring[k % 8])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        t0.record()
        for k in range(reps):
            env.step(ring[k % 8])
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1) * 1e3 / reps)
    return round(best, 3)


kw = dict(winning_score=15, serve="winner")
out["i32_us"] = steady(pikazoo_b200.PikaVecEnv(n, seed=2026, **kw), a32)
out["f32norm_us"] = steady(pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.float32, normalize_observation=True, **kw), a32)
out["f16norm_u8_us"] = steady(pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.float16, normalize_observation=True,
                                                       action_dtype=torch.uint8, **kw), a8)
out["bf16_fm_u8_us"] = steady(pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.bfloat16, normalize_observation=True,
                                                       action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40, **kw), a8)
out["ai_vs_ai_us"] = steady(pikazoo_b200.PikaVecEnv(n, seed=3, is_player1_computer=True, is_player2_computer=True, **kw), [None] * 8, 600)
print(json.dumps(out))
