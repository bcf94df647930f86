import os, sys, torch
sys.path.insert(0, os.getcwd())
import pikazoo_b200
from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout, rollout_fused
dev=torch.device("cuda",0)
world=8
total=8192*world+100
pol=MLPPolicy(device=dev, seed=3)
kw = dict(winning_score=5, serve="random", obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40)
for fused in (False, True):
    whole = pikazoo_b200.PikaVecEnv(total, device=dev, seed=7, **kw); whole.reset()
    if fused:
        for _ in range(4): rollout_fused(whole, pol, 32, seed=11)
    else:
        policy_rollout(whole, FusedActor(pol, whole, seed=11), 128)
    ws = whole.export_state()
    tot = torch.zeros_like(whole.stats)
    for rank in range(world):
        first,count = pikazoo_b200.shard_range(total, world, rank)
        mine = pikazoo_b200.make_sharded_env(total, rank, world, dev, seed=7, **kw); mine.reset()
        if fused:
            for _ in range(4): rollout_fused(mine, pol, 32, seed=11)
        else:
            policy_rollout(mine, FusedActor(pol, mine, seed=11), 128)
        eq = torch.equal(mine.export_state(), ws[first:first+count])
        if not eq:
            diff = (mine.export_state()!=ws[first:first+count]).any(1).nonzero().flatten()
            print("fused",fused,"rank",rank,"first",first,"count",count,"mismatching envs",diff.numel(), diff[:10].tolist())
        tot += mine.stats
    print("fused",fused,"stats equal", torch.equal(tot, whole.stats))
# part 1 of tests/multi_gpu_check.py: per-step path + K-frame rollouts, one computer player
kw = dict(winning_score=3, serve="random", is_player2_computer=True)
whole = pikazoo_b200.PikaVecEnv(total, device=dev, seed=99, **kw); whole.reset()
g = torch.Generator(device=dev).manual_seed(4242)
acts_all = [torch.randint(0, 18, (total, 2), generator=g, device=dev, dtype=torch.int32) for _ in range(96)]
for a in acts_all: whole.step(a)
for _ in range(3): whole.rollout(64, actions="synth", action_seed=5)
ws = whole.export_state(); tot = torch.zeros_like(whole.stats)
for rank in range(world):
    first,count = pikazoo_b200.shard_range(total, world, rank)
    mine = pikazoo_b200.make_sharded_env(total, rank, world, dev, seed=99, **kw); mine.reset()
    for a in acts_all: mine.step(a[first:first+count].clone())
    for _ in range(3): mine.rollout(64, actions="synth", action_seed=5)
    eq = torch.equal(mine.export_state(), ws[first:first+count])
    if not eq:
        diff = (mine.export_state()!=ws[first:first+count]).any(1).nonzero().flatten()
        print("part1 rank",rank,"first",first,"count",count,"mismatching envs",diff.numel(), diff[:10].tolist())
    tot += mine.stats
print("part1 stats equal", torch.equal(tot, whole.stats), (tot-whole.stats).tolist())
