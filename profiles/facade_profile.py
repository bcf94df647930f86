import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pikazoo_b200 import pikazoo_v0
env = pikazoo_v0.env(winning_score=15, seed=0)
env.reset()
v = env._vec
N=20000
t0=time.perf_counter()
for _ in range(N):
    v.step(env._actions); v.wait()
dt=time.perf_counter()-t0; print(f"vec.step+wait(spin): {dt/N*1e6:.1f} us")
t0=time.perf_counter()
for _ in range(N):
    v.step(env._actions); torch.cuda.current_stream().synchronize()
dt=time.perf_counter()-t0; print(f"vec.step+stream sync: {dt/N*1e6:.1f} us")
t0=time.perf_counter()
for _ in range(N):
    v.step(env._actions)
torch.cuda.synchronize()
dt=time.perf_counter()-t0; print(f"vec.step back to back (no wait): {dt/N*1e6:.1f} us")
# device-resident single env for comparison
import pikazoo_b200
d = pikazoo_b200.PikaVecEnv(1, seed=0); d.reset(); a=torch.zeros((1,2),dtype=torch.int32,device='cuda')
t0=time.perf_counter()
for _ in range(N):
    d.step(a)
torch.cuda.synchronize()
dt=time.perf_counter()-t0; print(f"device-resident vec.step back to back: {dt/N*1e6:.1f} us")
import cProfile, pstats
rng=np.random.default_rng(0); acts=rng.integers(0,18,size=(5000,2))
def run():
    for a in acts:
        if not env.agents: env.reset()
        env.step({"player_1": int(a[0]), "player_2": int(a[1])})
cProfile.run("run()", "/tmp/prof.out")
pstats.Stats("/tmp/prof.out").sort_stats("cumtime").print_stats(14)
