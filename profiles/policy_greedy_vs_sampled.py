"""ms per launch of pz_policy_mlp_act at 2 M envs, both implementations, sampled and greedy (what the sampler costs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pikazoo_b200 import _lib
from pikazoo_b200.policy import MLPPolicy
L = _lib.load()
n = 1 << 21
pol = MLPPolicy()
obs = torch.zeros(2, 40, n, dtype=torch.bfloat16, device="cuda")
obs[:, :35] = torch.rand(2, 35, n, device="cuda").bfloat16()
acts = torch.empty((n, 2), dtype=torch.uint8, device="cuda")
out = {}
for impl, name in ((1, "mma"), (0, "tc")):
    L.pz_policy_select(impl)
    for greedy in (False, True):
        for _ in range(5):
            pol.act_fused(obs, 0, out=acts, greedy=greedy)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for s in range(100):
            pol.act_fused(obs, s, out=acts, greedy=greedy)
        b.record(); torch.cuda.synchronize()
        out[f"{name}_greedy{int(greedy)}"] = a.elapsed_time(b) / 100
print(json.dumps(out))
