"""GPU: BASELINE.json's full sizes (>= 1 M envs per GPU) through size-independent properties —
the oracle cannot replay a million games in seconds, so these check invariants of the domain
plus exact agreement with the oracle on a strided sample of envs."""

import numpy as np
import pytest
import torch

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy

pytestmark = pytest.mark.gpu

N = 1 << 20


@pytest.fixture(scope="module")
def pz(cuda_lib):
    import pikazoo_b200

    return pikazoo_b200


def _obs_invariants(obs):
    # player blocks are swapped between agents, ball block shared (pikazoo_env.py:585-586)
    assert torch.equal(obs[:, 0, 0:13], obs[:, 1, 13:26]) and torch.equal(obs[:, 0, 13:26], obs[:, 1, 0:13])
    assert torch.equal(obs[:, 0, 26:35], obs[:, 1, 26:35])
    assert bool((obs[:, 0, 7:12].sum(1) == 1).all()) and bool((obs[:, 0, 20:25].sum(1) == 1).all())
    assert int(obs[:, 0, 0].min()) >= 32 and int(obs[:, 0, 0].max()) <= 184   # player 1 stays in its half
    assert int(obs[:, 0, 13].min()) >= 248 and int(obs[:, 0, 13].max()) <= 400
    assert int(obs[:, 0, 26].min()) >= 20 and int(obs[:, 0, 26].max()) <= 432  # ball x
    assert int(obs[:, 0, 27].min()) >= 0 and int(obs[:, 0, 27].max()) <= 252   # ball y


def test_one_million_envs_per_step_path(pz):
    steps, sample = 300, slice(0, N, 4099)
    idx = np.arange(N)[sample]
    env = pz.PikaVecEnv(N, seed=1234, winning_score=2, serve="random")
    orcs = po.OracleVecEnv(len(idx), seed=0, winning_score=2, serve="random")
    for j, i in enumerate(idx):  # oracle envs seeded like the sampled global envs
        po.lib().pk_init(po._p(orcs.state[j]), 1234 + int(i))
    obs = env.reset()
    assert np.array_equal(obs[sample].cpu().numpy(), orcs.reset())
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(steps):
        a = torch.randint(0, 18, (N, 2), generator=gen, device="cuda", dtype=torch.int32)
        obs, rew, done = env.step(a)
        o_obs, o_rew, o_done = orcs.step(a[sample].cpu().numpy())
        assert np.array_equal(obs[sample].cpu().numpy(), o_obs)
        assert np.array_equal(rew[sample].cpu().numpy(), o_rew.astype(np.float32))
        assert np.array_equal(done[sample].cpu().numpy(), o_done.astype(bool))
        if t % 50 == 0 or t == steps - 1:
            _obs_invariants(obs)
            assert torch.equal(rew[:, 0], -rew[:, 1]) and bool((rew.abs() <= 1).all())
            assert bool((rew[done][:, 0] != 0).all())  # a game can only end on a scoring frame
    d = env.stats_dict()
    assert d["calls"] == N * steps and d["p1_wins"] + d["p2_wins"] == d["episodes"]
    st = env.export_state()
    assert np.array_equal(st[sample].cpu().numpy(), orcs.state)
    assert int(st[:, 37:39].max()) <= 2


def test_one_million_envs_rollout_k64_ai_vs_ai(pz):
    # config 4: register-resident K = 64 rollout, computer vs computer
    cfg = dict(winning_score=15, serve="winner", is_player1_computer=True, is_player2_computer=True)
    env = pz.PikaVecEnv(N, seed=99, **cfg)
    env.reset()
    sample = slice(0, N, 8191)
    idx = np.arange(N)[sample]
    orcs = po.OracleVecEnv(len(idx), seed=0, **cfg)
    for j, i in enumerate(idx):
        po.lib().pk_init(po._p(orcs.state[j]), 99 + int(i))
    orcs.reset()
    for l in range(4):
        env.rollout(64)
        orcs.rollout(64)
    st = env.export_state()
    assert np.array_equal(st[sample].cpu().numpy(), orcs.state)
    # determinism + equivalence of K = 64 launches and 256 per-step launches on a second instance
    env2 = pz.PikaVecEnv(N, seed=99, **cfg)
    env2.reset()
    for _ in range(256):
        env2.step(None)
    assert torch.equal(env2.state, env.state)


def test_sharding_is_invisible(pz):
    # 8(e): trajectories are invariant to how the global batch is split over ranks
    total, steps = 300_000, 120
    whole = pz.make_sharded_env(total, 0, 1, "cuda", seed=7, winning_score=1, serve="random")
    parts = [pz.make_sharded_env(total, r, 3, "cuda", seed=7, winning_score=1, serve="random") for r in range(3)]
    o = whole.reset()
    assert torch.equal(o, torch.cat([p.reset() for p in parts]))
    for t in range(steps):
        a = torch.from_numpy(synth_actions_numpy(3, 0, total, t, 18)).cuda()
        o, r, d = whole.step(a)
        outs = [p.step(a[p.first_env:p.first_env + p.num_envs].contiguous()) for p in parts]
        assert torch.equal(o, torch.cat([x[0] for x in outs]))
        assert torch.equal(d, torch.cat([x[2] for x in outs]))
    tot = torch.stack([p.stats for p in parts]).sum(0)
    assert torch.equal(tot, whole.stats)


def test_sixteen_million_envs_one_device(pz):
    """configs[4]'s total batch (16,777,216 envs, serve='random', winning_score=5) on a single device:
    64-bit indexing of the 4.7 GB observation tensor, exact agreement with the oracle on envs sampled
    from the whole index range (including the last warp), in both observation layouts."""
    n, steps = 1 << 24, 40
    cfg = dict(winning_score=5, serve="random")
    idx = np.concatenate([np.arange(0, 64), np.arange(n // 2 - 32, n // 2 + 32), np.arange(n - 64, n),
                          np.arange(0, n, 1_000_003)])
    tidx = torch.from_numpy(idx).cuda()
    orcs = po.OracleVecEnv(len(idx), seed=0, **cfg)
    for j, i in enumerate(idx):
        po.lib().pk_init(po._p(orcs.state[j]), 77 + int(i))
    env = pz.PikaVecEnv(n, seed=77, obs_dtype=torch.int16, action_dtype=torch.uint8, **cfg)
    fm = pz.PikaVecEnv(n, seed=77, obs_dtype=torch.int16, action_dtype=torch.uint8, obs_layout="feature_major", **cfg)
    assert np.array_equal(env.reset()[tidx].cpu().numpy(), orcs.reset().astype(np.int16))
    fm.reset()
    gen = torch.Generator(device="cuda").manual_seed(11)
    for t in range(steps):
        a = torch.randint(0, 18, (n, 2), generator=gen, device="cuda", dtype=torch.uint8)
        obs, rew, done = env.step(a)
        fobs, frew, fdone = fm.step(a)
        o_obs, o_rew, o_done = orcs.step(a[tidx].cpu().numpy())
        assert np.array_equal(obs[tidx].cpu().numpy(), o_obs.astype(np.int16)), t
        assert np.array_equal(fobs[:, :, tidx].permute(2, 0, 1).cpu().numpy(), o_obs.astype(np.int16)), t
        assert np.array_equal(rew[tidx].cpu().numpy(), o_rew.astype(np.float32))
        assert torch.equal(done, fdone) and torch.equal(rew, frew)
    assert torch.equal(env.state, fm.state)
    assert env.stats_dict()["calls"] == n * steps
    del env, fm
    torch.cuda.empty_cache()


def test_calls_follow_the_current_stream(pz):
    """every entry point is asynchronous on torch's current stream: two envs stepped on two side streams,
    interleaved, must equal the same envs stepped on the default stream"""
    n, steps = 50_000, 60
    cfg = dict(winning_score=3, serve="random", is_player1_computer=True)
    ref = [pz.PikaVecEnv(n, seed=s, **cfg) for s in (1, 2)]
    side = [pz.PikaVecEnv(n, seed=s, **cfg) for s in (1, 2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    acts = [torch.from_numpy(synth_actions_numpy(3, 0, n, t, 18)).cuda() for t in range(steps)]
    torch.cuda.synchronize()
    for e in ref:
        e.reset()
        for t in range(steps):
            e.step(acts[t])
    for e, s in zip(side, streams):
        with torch.cuda.stream(s):
            e.reset()
    for t in range(steps):
        for e, s in zip(side, streams):
            with torch.cuda.stream(s):
                e.step(acts[t])
    torch.cuda.synchronize()
    for a, b in zip(ref, side):
        assert torch.equal(a.state, b.state) and torch.equal(a.obs, b.obs) and torch.equal(a.stats, b.stats)


def test_configs2_at_its_stated_size(pz):
    """configs[2] as BASELINE.json states it: 65,536 envs with SimplifyAction + RewardByBallPosition fused into the
    step kernel — every output compared with the oracle on a strided sample of envs (every 61st: 1,075 envs incl.
    the last warp), invariants on the whole batch, 600 frames."""
    n, steps = 65_536, 600
    shaped = ((0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4), 216, 176)
    cfg = dict(winning_score=15, serve="winner", simplify_action=True, reward_by_ball_position=shaped)
    idx = np.unique(np.concatenate([np.arange(0, n, 61), np.arange(n - 40, n)]))
    tidx = torch.from_numpy(idx).cuda()
    env = pz.PikaVecEnv(n, seed=4321, reward_dtype=torch.float64, **cfg)
    orc = po.OracleVecEnv(len(idx), seeds=4321 + idx, **cfg)
    assert np.array_equal(env.reset()[tidx].cpu().numpy(), orc.reset())
    gen = torch.Generator(device="cuda").manual_seed(3)
    zones = set()
    for t in range(steps):
        a = torch.randint(0, 13, (n, 2), generator=gen, device="cuda", dtype=torch.int32)
        obs, rew, done = env.step(a)
        o_obs, o_rew, o_done = orc.step(a[tidx].cpu().numpy())
        assert np.array_equal(obs[tidx].cpu().numpy(), o_obs), t
        assert np.array_equal(rew[tidx].cpu().numpy(), o_rew), t   # float64 rewards: Python's int + float, bit for bit
        assert np.array_equal(done[tidx].cpu().numpy(), o_done.astype(bool)), t
        if t % 100 == 0:
            _obs_invariants(obs)
            zones |= set(torch.unique((rew[:, 0] * 10).round()).tolist())
    assert np.array_equal(env.export_state()[tidx].cpu().numpy(), orc.state)
    assert len(zones) >= 4 and env.stats_dict()["bad_actions"] == 0


def test_configs4_two_kernel_loop_at_its_stated_size(pz):
    """configs[4] per GPU as stated: 2,097,152 envs, winning_score 5, serve random, the MLP policy in the loop through
    the two-kernel path (pz_policy_mlp_act -> pz_step on feature-major bf16 observations); the oracle replays the
    sampled actions of a strided sample of envs."""
    from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout

    n, steps = 1 << 21, 150
    cfg = dict(winning_score=5, serve="random")
    env = pz.PikaVecEnv(n, seed=606, obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.uint8,
                        obs_layout="feature_major", obs_feature_rows=40, **cfg)
    idx = np.unique(np.concatenate([np.arange(0, n, 4099), np.arange(n - 33, n)]))
    tidx = torch.from_numpy(idx).cuda()
    orc = po.OracleVecEnv(len(idx), seeds=606 + idx, **cfg)
    env.reset(), orc.reset()
    actor = FusedActor(MLPPolicy(device=env.device, seed=3), env, seed=12)

    def mirror(t, actions, obs, reward, done):
        orc.step(actions[tidx].cpu().numpy().astype(np.int32))
        assert np.array_equal(done[tidx].cpu().numpy(), orc.done.astype(bool)), t
        if t % 50 == 0 or t == steps - 1:
            got = obs[:, :35, :][:, :, tidx].permute(2, 0, 1).contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
            assert np.array_equal(got, orc.normalized_obs("bfloat16")), t

    policy_rollout(env, actor, steps, on_step=mirror)
    assert np.array_equal(env.export_state()[tidx].cpu().numpy(), orc.state)
    st = env.stats_dict()
    assert st["calls"] == n * steps and st["episodes"] > 0 and st["p1_wins"] + st["p2_wins"] == st["episodes"]
