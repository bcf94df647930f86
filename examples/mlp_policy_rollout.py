#!/usr/bin/env python
"""configs[4]: envs sharded over the GPUs of one box, serve='random', winning_score=5, actions from an
on-device MLP policy in a rollout loop; NCCL all-reduces the episode statistics every M steps.

    python examples/mlp_policy_rollout.py --envs-per-gpu 2097152 --steps 200
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/mlp_policy_rollout.py --envs-per-gpu 2097152 --steps 200
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pikazoo_b200  # noqa: E402
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout, rollout_fused  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--stats-every", type=int, default=50)
    ap.add_argument("--actor", choices=["rollout", "fused", "eager"], default="rollout",
                    help="rollout: the whole loop in one launch per K frames (pz_rollout_policy: env, observation tile, "
                         "both layers on tcgen05 and the sample all on chip); fused: two launches per frame "
                         "(pz_policy_mlp_act -> pz_step); eager: MLPPolicy.act in PyTorch -> pz_step")
    ap.add_argument("--K", type=int, default=50, help="frames per launch of the rollout actor")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = pikazoo_b200.make_sharded_env(a.envs_per_gpu * world, rank, world, dev, seed=5, winning_score=5,
                                        serve="random", obs_dtype=torch.bfloat16, normalize_observation=True,
                                        action_dtype=torch.int64 if a.actor == "eager" else torch.uint8,
                                        obs_layout="feature_major", obs_feature_rows=40)
    policy = MLPPolicy(device=dev)
    act = FusedActor(policy, env, seed=1) if a.actor == "fused" else policy.act

    def run(frames):
        if a.actor == "rollout":
            for f0 in range(0, frames, a.K):
                rollout_fused(env, policy, min(a.K, frames - f0), seed=1)
        else:
            policy_rollout(env, act, frames)

    env.reset()
    run(10)
    pikazoo_b200.allreduce_stats(env.stats.clone())  # NCCL communicator set-up happens on the first collective
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done_steps = 0
    while done_steps < a.steps:
        k = min(a.stats_every, a.steps - done_steps)
        run(k)
        done_steps += k
        stats = env.stats.clone()
        pikazoo_b200.allreduce_stats(stats)  # the one collective: 16 int64 over NVLink
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        names = pikazoo_b200._lib.STAT_NAMES
        print({n: int(stats[i]) for i, n in enumerate(names)})
        print(f"{a.envs_per_gpu * world * a.steps / dt / 1e9:.2f} G env-steps/s over {world} GPU(s), {a.actor} policy in the loop")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
