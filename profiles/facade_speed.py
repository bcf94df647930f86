import sys, os, time
sys.path.insert(0, os.getcwd())
from pikazoo_b200 import pikazoo_v0
env = pikazoo_v0.env(winning_score=15, seed=0)
env.reset()
import numpy as np
rng = np.random.default_rng(0)
acts = rng.integers(0, 18, size=(20000, 2))
t0 = time.perf_counter(); n = 0
for a in acts:
    if not env.agents: env.reset()
    env.step({"player_1": int(a[0]), "player_2": int(a[1])}); n += 1
dt = time.perf_counter() - t0
print(f"facade: {n/dt:.0f} steps/s ({dt/n*1e6:.1f} us per step)")
