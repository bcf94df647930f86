import sys, os
sys.path.insert(0, os.getcwd())
import torch, pikazoo_b200
from pikazoo_b200.policy import MLPPolicy
n=1<<21
env=pikazoo_b200.PikaVecEnv(n, seed=5, winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.int64, obs_layout="feature_major", obs_feature_rows=40)
pol=MLPPolicy()
obs=env.reset()
for _ in range(5): a=pol.act(obs); env.step(a)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        a=pol.act(obs); env.step(a)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
