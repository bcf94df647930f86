"""Run under torchrun with one rank per GPU (tests/test_gpu_multi.py launches it; bench.py --gpus N runs the same
check before its timed region): every rank steps ITS shard of a global batch and also the whole batch alone; its
shard of the whole-batch state must equal its own state bit for bit, the NCCL all-reduced statistics must equal the
whole-batch statistics, and — for the MLP-policy loop of configs[4] — the sampled actions must not depend on how the
batch is sharded either (the policy's counter stream is keyed by the global env index)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout, rollout_fused  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    total = 8192 * world + 100  # ragged shards
    first, count = pikazoo_b200.shard_range(total, world, rank)
    ok = True
    # 1. per-step path + K-frame rollouts, one computer player
    kw = dict(winning_score=3, serve="random", is_player2_computer=True)
    mine = pikazoo_b200.make_sharded_env(total, rank, world, dev, seed=99, **kw)
    whole = pikazoo_b200.PikaVecEnv(total, device=dev, seed=99, **kw)
    g = torch.Generator(device=dev).manual_seed(4242)  # same seed on every rank: the same global action tensor
    mine.reset(), whole.reset()
    for _ in range(96):
        acts = torch.randint(0, 18, (total, 2), generator=g, device=dev, dtype=torch.int32)
        mine.step(acts[first:first + count].clone())  # (a fresh tensor: the library wants 16-byte aligned actions)
        whole.step(acts)
    for _ in range(3):
        mine.rollout(64, actions="synth", action_seed=5)
        whole.rollout(64, actions="synth", action_seed=5)
    ok &= bool(torch.equal(mine.export_state(), whole.export_state()[first:first + count]))
    summed = mine.stats.clone()
    pikazoo_b200.allreduce_stats(summed)
    ok &= bool(torch.equal(summed, whole.stats))
    # 2. configs[4]: the policy in the loop, both as two kernels per frame and as one launch per K frames
    kw = dict(winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True,
              action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40)
    pol = MLPPolicy(device=dev, seed=3)
    for fused in (False, True):
        mine = pikazoo_b200.make_sharded_env(total, rank, world, dev, seed=7, **kw)
        whole = pikazoo_b200.PikaVecEnv(total, device=dev, seed=7, **kw)
        mine.reset(), whole.reset()
        if fused:
            for _ in range(4):
                rollout_fused(mine, pol, 32, seed=11)
                rollout_fused(whole, pol, 32, seed=11)
        else:
            policy_rollout(mine, FusedActor(pol, mine, seed=11), 128)
            policy_rollout(whole, FusedActor(pol, whole, seed=11), 128)
        ok &= bool(torch.equal(mine.export_state(), whole.export_state()[first:first + count]))
        summed = mine.stats.clone()
        pikazoo_b200.allreduce_stats(summed)
        ok &= bool(torch.equal(summed, whole.stats))
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if int(flag.item()) else "FAILED", "ranks", world, "global envs", total, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
