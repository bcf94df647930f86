"""Host-side plumbing for configs[4] of BASELINE.json: actions for both agents from an on-device torch
MLP policy inside a rollout loop. The policy is ordinary PyTorch (cuBLAS GEMMs, not the product); what
it demonstrates is that the simulator's tensors feed a policy and take its sampled actions without ever
leaving the device: observations arrive as normalised fp16/bf16 rows straight from the step kernel
(NormalizeObservation fused), actions return as the int64 tensor torch's argmax produces."""

from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from .vec_env import PikaVecEnv


class MLPPolicy(nn.Module):
    """obs [*, 35] -> logits [*, n_actions]; one set of weights per agent (self-play shares none)."""

    def __init__(self, n_actions: int = 18, hidden: int = 64, dtype: torch.dtype = torch.bfloat16, device="cuda",
                 seed: int = 0):
        super().__init__()
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.n_actions = n_actions
        # [agent, in, out] so both agents run as one batched matmul
        self.w1 = nn.Parameter((torch.randn(2, 35, hidden, generator=g) / 35 ** 0.5).to(device=device, dtype=dtype))
        self.b1 = nn.Parameter(torch.zeros(2, 1, hidden, device=device, dtype=dtype))
        self.w2 = nn.Parameter((torch.randn(2, hidden, n_actions, generator=g) / hidden ** 0.5).to(device=device, dtype=dtype))
        self.b2 = nn.Parameter(torch.zeros(2, 1, n_actions, device=device, dtype=dtype))

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        """obs [N, 2, 35] (any float dtype) -> logits [N, 2, n_actions]."""
        x = obs.to(self.w1.dtype).transpose(0, 1)                 # [2, N, 35]
        h = torch.relu(torch.baddbmm(self.b1, x, self.w1))        # [2, N, hidden]
        return torch.baddbmm(self.b2, h, self.w2).transpose(0, 1)  # [N, 2, n_actions]

    @torch.no_grad()
    def act(self, obs: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """Categorical sample by Gumbel-max: int64 actions [N, 2]."""
        logits = self.forward(obs).float()
        u = torch.rand(logits.shape, device=logits.device, generator=generator).clamp_(min=1e-20)
        return (logits - torch.log(-torch.log(u))).argmax(dim=-1)


@torch.no_grad()
def policy_rollout(env: PikaVecEnv, policy: Callable[[torch.Tensor], torch.Tensor], steps: int,
                   on_step: Optional[Callable] = None) -> torch.Tensor:
    """`steps` iterations of obs -> policy -> env.step on the device. `policy(obs)` returns actions
    [N, 2] of env.action_dtype. Returns the last observation. `on_step(t, actions, obs, reward, done)` is
    called after every step (tests use it to mirror the run on the oracle)."""
    obs = env.obs
    for t in range(steps):
        actions = policy(obs)
        obs, reward, done = env.step(actions)
        if on_step is not None:
            on_step(t, actions, obs, reward, done)
    return obs
