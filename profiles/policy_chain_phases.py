"""Cycle stamps of one chain of the tcgen05 policy kernel (clock64 at its phase boundaries, iterations 4..11 of chain
0 of CTA 0), from a debug build that writes them into the logits buffer:

    PZ_NVCC_FLAGS=-DPZ_TC_TIMING python pikazoo_b200/build.py --out variants/timing.so
    PIKAZOO_B200_LIB=variants/timing.so python profiles/policy_chain_phases.py
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pikazoo_b200 import _lib
from pikazoo_b200.policy import MLPPolicy
L = _lib.load(); L.pz_policy_select(0)
n = 1 << 21
pol = MLPPolicy()
obs = torch.zeros(2, 40, n, dtype=torch.bfloat16, device="cuda")
obs[:, :35] = torch.rand(2, 35, n, device="cuda").bfloat16()
acts = torch.empty((n, 2), dtype=torch.uint8, device="cuda")
lg = torch.zeros((n, 2, 18), device="cuda")
for greedy in (False, True):
    for _ in range(3):
        pol.act_fused(obs, 0, out=acts, greedy=greedy, logits_out=lg)
    torch.cuda.synchronize()
    t = lg.view(-1)[: 8 * 12 * 2].view(torch.int64).cpu().view(8, 12)[:, :11]
    d = (t[:, 1:] - t[:, :-1]).tolist()
    print("greedy" if greedy else "sampled")
    print(" wait1 epi1 sync mma2 pre+noise wait2 ldlog ready mma1 sample")
    for row in d: print(row)
    print("iter period:", (t[1:, 0] - t[:-1, 0]).tolist())
