"""ncu driver: a few launches of the feature-major bf16 normalised step kernel at 1,048,576 envs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = torch.Generator(device="cuda").manual_seed(1)
a8 = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32).to(torch.uint8) for _ in range(2)]
e = pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.uint8,
                            obs_layout="feature_major", obs_feature_rows=40, winning_score=15)
e.reset()
for k in range(6):
    e.step(a8[k % 2])
torch.cuda.synchronize()
print("ok")
