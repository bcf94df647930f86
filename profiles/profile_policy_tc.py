"""A few launches of pz_policy_mlp_act (tcgen05 implementation unless PZ_IMPL=1) at 2 M envs, for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pikazoo_b200 import _lib  # noqa: E402
from pikazoo_b200.policy import MLPPolicy  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
_lib.load().pz_policy_select(int(os.environ.get("PZ_IMPL", "0")))
pol = MLPPolicy()
obs = torch.zeros(2, 40, n, dtype=torch.bfloat16, device="cuda")
obs[:, :35] = torch.rand(2, 35, n, device="cuda").bfloat16()
acts = torch.empty((n, 2), dtype=torch.uint8, device="cuda")
for s in range(4):
    pol.act_fused(obs, s, out=acts)
torch.cuda.synchronize()
print("ok")
