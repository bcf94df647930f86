"""ncu driver: a few launches of each kernel variant at 1,048,576 envs (see profiles/README in DESIGN.md §6).
   order of launches: 4x step<0,i32>  4x step<0,f16 normalised,u8 actions>  4x step<3,i32> (computer vs computer)
                      4x rollout<3> K=64"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pikazoo_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kw = dict(winning_score=15, serve="winner")
g = torch.Generator(device="cuda").manual_seed(1)
a32 = [torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32) for _ in range(2)]
a8 = [a.to(torch.uint8) for a in a32]

e = pikazoo_b200.PikaVecEnv(n, seed=1, **kw)
e.reset()
for k in range(reps):
    e.step(a32[k % 2])
torch.cuda.synchronize()
del e
e = pikazoo_b200.PikaVecEnv(n, seed=2, obs_dtype=torch.float16, normalize_observation=True, action_dtype=torch.uint8, **kw)
e.reset()
for k in range(reps):
    e.step(a8[k % 2])
torch.cuda.synchronize()
del e
e = pikazoo_b200.PikaVecEnv(n, seed=3, is_player1_computer=True, is_player2_computer=True, **kw)
e.reset()
e.rollout(64)  # move away from the serve position so the frames profiled are mid-rally
for k in range(reps):
    e.step(None)
torch.cuda.synchronize()
for k in range(reps):
    e.rollout(64)
torch.cuda.synchronize()
print("ok", e.stats_dict())
