"""Multi-GPU plumbing: envs shard across ranks by global env index with NO communication on
the step path; the only collective is an all-reduce(SUM) of the episode
statistics vector (NCCL over NVLink on GPUs, gloo in the CPU tests)."""

from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[first, first + count) of the global env indices owned by `rank` (contiguous, balanced)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def make_sharded_env(total_envs: int, rank: int, world_size: int, device, seed: int = 0, **kwargs):
    """The rank's shard as a PikaVecEnv. Env i of the global batch draws from
    PCG64(seed + i) whichever rank owns it, so trajectories do not depend on the GPU count."""
    from .vec_env import PikaVecEnv

    first, count = shard_range(total_envs, world_size, rank)
    return PikaVecEnv(count, device=device, seed=seed, first_env=first, **kwargs)


def allreduce_stats(stats: torch.Tensor, group=None, async_op: bool = False):
    """SUM the int64 statistics vector over ranks, in place. Returns the work handle if async."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
