"""The tcgen05 implementation of pz_policy_mlp_act against the warp-level one and a PyTorch fp32 reference:
logits, greedy and sampled actions at several batch sizes (incl. ragged tiles), then ms per launch of both at 2 M envs.

    python profiles/check_policy_tc.py [OUT.json]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pikazoo_b200 import _lib  # noqa: E402
from pikazoo_b200.policy import MLPPolicy, inverse_cdf_reference  # noqa: E402

TC, MMA = 0, 1
L = _lib.load()
out = {"cases": []}


def run(pol, obs, impl, step, greedy=False, want_logits=True):
    prev = L.pz_policy_select(impl)
    try:
        n = obs.shape[2]
        lg = torch.full((n, 2, pol.n_actions), float("nan"), device="cuda") if want_logits else None
        act = pol.act_fused(obs, step, seed=3, first_env=11, greedy=greedy, logits_out=lg).clone()
        torch.cuda.synchronize()
        return act, lg
    finally:
        L.pz_policy_select(prev)


def reference_logits(pol, obs):
    x = obs.float()                                              # [2, 40, N]
    h = torch.bmm(pol.w1.float(), x).relu_().bfloat16().float()  # hidden activations rounded to bf16
    return torch.bmm(pol.w2.float(), h).permute(2, 0, 1)         # [N, 2, A]


ok = True
for n_actions in (18, 13, 7):
    pol = MLPPolicy(n_actions=n_actions, seed=1)
    with torch.no_grad():  # non-trivial biases
        pol.w1[:, :pol.hidden, pol.ONES_ROW] = torch.randn(2, pol.hidden, device="cuda").bfloat16() * 0.3
        pol.w2[:, :, pol.hidden] = torch.randn(2, n_actions, device="cuda").bfloat16() * 0.3
    for n in (128, 1000, 4096 + 8, 1 << 16):
        g = torch.Generator(device="cuda").manual_seed(n)
        obs = torch.zeros(2, 40, n, dtype=torch.bfloat16, device="cuda")
        obs[:, :35] = torch.rand(2, 35, n, generator=g, device="cuda").bfloat16()
        a_m, l_m = run(pol, obs, MMA, 5)
        a_t, l_t = run(pol, obs, TC, 5)
        ref = reference_logits(pol, obs)
        g_m, _ = run(pol, obs, MMA, 5, greedy=True)
        g_t, _ = run(pol, obs, TC, 5, greedy=True)
        case = {
            "n_actions": n_actions, "n": n,
            "logits_tc_vs_ref": float((l_t - ref).abs().max()),
            "logits_mma_vs_ref": float((l_m - ref).abs().max()),
            "logits_tc_vs_mma": float((l_t - l_m).abs().max()),
            "sampled_mismatch": float((a_t.cpu().numpy().astype("int64")
                                       != inverse_cdf_reference(l_t.cpu().numpy(), 3, 5, 11)).mean()),
            "greedy_mismatch": float((g_t != g_m).float().mean()),
            "nan_logits_tc": int(torch.isnan(l_t).sum()),
        }
        case["ok"] = (case["logits_tc_vs_ref"] < 1e-2 and case["logits_tc_vs_mma"] < 1e-4 and case["sampled_mismatch"] < 1e-3
                      and case["greedy_mismatch"] < 1e-3 and case["nan_logits_tc"] == 0)
        ok &= case["ok"]
        out["cases"].append(case)
out["ok"] = bool(ok)

n = 1 << 21
pol = MLPPolicy()
obs = torch.zeros(2, 40, n, dtype=torch.bfloat16, device="cuda")
obs[:, :35] = torch.rand(2, 35, n, device="cuda").bfloat16()
acts = torch.empty((n, 2), dtype=torch.uint8, device="cuda")
for impl, name in ((MMA, "mma_sync"), (TC, "tcgen05")):
    L.pz_policy_select(impl)
    for _ in range(5):
        pol.act_fused(obs, 0, out=acts)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for s in range(100):
        pol.act_fused(obs, s, out=acts)
    b.record()
    torch.cuda.synchronize()
    out[name + "_ms_per_2M_envs"] = a.elapsed_time(b) / 100
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
sys.exit(0 if ok else 1)
