"""ncu driver for the bench's main kernel in STEADY STATE: 1,048,576 envs advanced 2,048 frames (as bench.py does), then a
run of per-step launches. Profile a launch in the middle with `--cache-control none` to see the memory traffic the
kernel has between its neighbours (state lines of the previous launch still in L2), or with the default cache flush
for the cold-cache figure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = torch.Generator(device="cuda").manual_seed(1)
ring = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(8)]
e = pikazoo_b200.PikaVecEnv(n, seed=2026, winning_score=15, serve="winner")
e.reset()
for _ in range(8):
    e.rollout(256, actions="synth", action_seed=77)
for k in range(launches):
    e.step(ring[k % 8])
torch.cuda.synchronize()
print("ok", e.stats_dict())
