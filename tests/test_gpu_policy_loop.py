"""GPU: configs[4] of BASELINE.json in miniature — serve='random', winning_score=5, both agents' actions
sampled on the device from a torch MLP fed by the kernel's normalised bf16 observations, int64 actions
straight from argmax. The oracle replays the very actions the policy chose, so observations (as bf16 bit
patterns), rewards, dones and the full hidden state must agree exactly."""

import numpy as np
import pytest
import torch

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def test_mlp_policy_rollout_loop_matches_oracle(cuda_lib):
    import pikazoo_b200
    from pikazoo_b200.policy import MLPPolicy, policy_rollout

    n, steps = 8192, 600
    cfg = dict(winning_score=5, serve="random")
    env = pikazoo_b200.PikaVecEnv(n, seed=16, obs_dtype=torch.bfloat16, normalize_observation=True,
                                  action_dtype=torch.int64, reward_dtype=torch.float64, **cfg)
    orc = po.OracleVecEnv(n, seed=16, **cfg)
    policy = MLPPolicy(device=env.device, seed=3)
    gen = torch.Generator(device=env.device).manual_seed(99)
    obs0 = env.reset()
    orc.reset()
    assert np.array_equal(obs0.view(torch.int16).cpu().numpy().view(np.uint16), orc.normalized_obs("bfloat16"))
    seen = {"done": 0, "distinct": set()}

    def mirror(t, actions, obs, reward, done):
        a = actions.cpu().numpy()
        assert a.dtype == np.int64 and a.min() >= 0 and a.max() < 18
        orc.step(a.astype(np.int32))
        if t % 25 == 0 or t == steps - 1:
            assert np.array_equal(obs.view(torch.int16).cpu().numpy().view(np.uint16), orc.normalized_obs("bfloat16")), t
            assert np.array_equal(reward.cpu().numpy(), orc.reward)
        assert np.array_equal(done.cpu().numpy(), orc.done.astype(bool)), t
        seen["done"] += int(done.sum())
        seen["distinct"] |= set(np.unique(a).tolist())

    policy_rollout(env, lambda o: policy.act(o, gen), steps, on_step=mirror)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    assert seen["done"] > n // 2 and len(seen["distinct"]) == 18  # games finish; the policy uses every action


def test_policy_consumes_feature_major_observations(cuda_lib):
    """Same loop with obs_layout='feature_major' ([2, 40, N] bf16, the GEMM operand as it is): the policy
    must produce the same logits from either layout, and the run must match the oracle."""
    import pikazoo_b200
    from pikazoo_b200.policy import MLPPolicy, policy_rollout

    n, steps = 4096, 300
    cfg = dict(winning_score=5, serve="random")
    kw = dict(seed=16, obs_dtype=torch.bfloat16, normalize_observation=True, action_dtype=torch.int64, **cfg)
    fm = pikazoo_b200.PikaVecEnv(n, obs_layout="feature_major", obs_feature_rows=40, **kw)
    em = pikazoo_b200.PikaVecEnv(n, **kw)
    orc = po.OracleVecEnv(n, seed=16, **cfg)
    policy = MLPPolicy(device=fm.device, seed=3)
    fm.reset(), em.reset(), orc.reset()
    assert torch.equal(policy.logits_t(fm.obs), policy.logits_t(em.obs))
    gen = torch.Generator(device=fm.device).manual_seed(7)

    def mirror(t, actions, obs, reward, done):
        orc.step(actions.cpu().numpy().astype(np.int32))
        em.step(actions)
        if t % 50 == 0 or t == steps - 1:
            assert torch.equal(obs[:, :35, :], em.obs.permute(1, 2, 0))
            assert np.array_equal(done.cpu().numpy(), orc.done.astype(bool))

    policy_rollout(fm, lambda o: policy.act(o, gen), steps, on_step=mirror)
    assert np.array_equal(fm.export_state().cpu().numpy(), orc.state)


def test_step_loop_is_cuda_graph_capturable(cuda_lib):
    """The launch-bound small-batch regime (configs[1], 4,096 envs): K steps with on-device action
    sampling captured once in a CUDA graph and replayed must equal the same K steps issued eagerly."""
    import pikazoo_b200

    n, K, replays = 4096, 16, 20
    kw = dict(seed=21, winning_score=5, serve="random", is_player2_computer=True, action_dtype=torch.int64)
    eager = pikazoo_b200.PikaVecEnv(n, **kw)
    graphed = pikazoo_b200.PikaVecEnv(n, **kw)
    eager.reset(), graphed.reset()
    assert cuda_lib.pz_tables_prepare(None) == 0  # table construction synchronises: do it before capturing
    # counter-based on-device action source (a stand-in policy that needs no RNG state inside the graph)
    frame_e = torch.zeros((), dtype=torch.int64, device="cuda")
    frame_g = torch.zeros((), dtype=torch.int64, device="cuda")
    idx = torch.arange(2 * n, device="cuda", dtype=torch.int64).view(n, 2)

    def act(frame):
        return ((idx * 2654435761 + frame * 40503) >> 7) % 18

    def k_steps(env, frame):
        for _ in range(K):
            env.step(act(frame))
            frame += 1

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        k_steps(graphed, frame_g)  # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    k_steps(eager, frame_e)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        k_steps(graphed, frame_g)
    k_steps(eager, frame_e)  # the capture itself does not execute: replay once to catch up
    g.replay()
    for _ in range(replays):
        g.replay()
        k_steps(eager, frame_e)
    torch.cuda.synchronize()
    assert int(frame_g) == int(frame_e) == K * (replays + 2)
    assert torch.equal(graphed.export_state(), eager.export_state())
    assert torch.equal(graphed.obs, eager.obs) and torch.equal(graphed.stats, eager.stats)
