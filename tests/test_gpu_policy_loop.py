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


# ---- the fused policy kernel (csrc/pz_policy_tc.cu: tcgen05 + TMEM, the default; csrc/pz_policy.cu: mma.sync) ----
IMPLS = {"tcgen05": 0, "mma_sync": 1}


@pytest.fixture(params=sorted(IMPLS))
def fused_impl(request, cuda_lib):
    """Every test of the fused kernel runs on both implementations of pz_policy_mlp_act."""
    prev = cuda_lib.pz_policy_select(IMPLS[request.param])
    assert prev in (0, 1)
    yield request.param
    cuda_lib.pz_policy_select(prev)


def _reference_logits(policy, obs):
    """[N, 2, A] float32 from plain PyTorch fp32 arithmetic on the same bf16 values, hidden activations rounded
    to bf16 where the kernel rounds them."""
    x = obs[:, : policy.K_PAD, :].float()                                 # [2, 40, N]
    h = torch.bmm(policy.w1.float(), x).to(torch.bfloat16).float().relu_()  # [2, 72, N]
    return torch.bmm(policy.w2.float(), h).permute(2, 0, 1).contiguous()


def _expected_samples(impl, logits, seed, step, first_env):
    """What the implementation's categorical sampler must return for these logits and counters (numpy restatements
    in pikazoo_b200/policy.py): inversion of the cumulative distribution (tcgen05), Gumbel arg-max (mma.sync)."""
    from pikazoo_b200.policy import gumbel_noise_reference, inverse_cdf_reference, sample_reference

    n, _, n_actions = logits.shape
    if impl == "tcgen05":
        return inverse_cdf_reference(logits, seed, step, first_env)
    return sample_reference(logits, gumbel_noise_reference(seed, step, first_env, n, n_actions))


def _played_env(n, **kw):
    import pikazoo_b200

    env = pikazoo_b200.PikaVecEnv(n, seed=31, winning_score=5, serve="random", obs_dtype=torch.bfloat16,
                                  normalize_observation=True, obs_layout="feature_major", obs_feature_rows=40, **kw)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(40):  # spread the envs over different states
        env.step(torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=env.action_dtype))
    return env


@pytest.mark.parametrize("n", [64, 1000, 4096 + 8, 100_003])
def test_fused_policy_logits_and_greedy_actions(cuda_lib, fused_impl, n):
    """Floating point: the fused kernel's logits against a PyTorch fp32 reference of the same network.
    Tolerance 2e-3 absolute (fp32 sums in another order can round a hidden activation to the neighbouring
    bf16, 2^-8 relative, which reaches a logit through one weight of magnitude ~0.1). Greedy actions are the
    arg-max of the kernel's own logits, exactly. Sizes cover full tiles, the ragged last tile and the
    unaligned-row path (n % 8 != 0)."""
    from pikazoo_b200.policy import MLPPolicy, sample_reference

    env = _played_env(n)
    policy = MLPPolicy(device=env.device, seed=3)
    logits = torch.full((n, 2, 18), float("nan"), device="cuda")
    for dtype in (torch.uint8, torch.int32, torch.int64):
        a = policy.act_fused(env.obs, step=0, action_dtype=dtype, greedy=True, logits_out=logits)
        assert a.dtype == dtype and a.shape == (n, 2)
        assert np.array_equal(a.cpu().numpy().astype(np.int64), sample_reference(logits.cpu().numpy(), None))
        # i.e. the arg-max of the logits, except between logits equal in their upper 27 bits
        differ = a.long() != logits.argmax(dim=2)
        if bool(differ.any()):
            top2 = logits[differ].topk(2, dim=-1).values
            assert float((top2[:, 0] - top2[:, 1]).abs().max()) < 1e-4 and int(differ.sum()) < 5
    ref = _reference_logits(policy, env.obs)
    assert torch.isfinite(logits).all()
    assert (logits - ref).abs().max().item() < 2e-3
    # and the eager policy (bf16 logits) agrees to bf16 resolution
    eager = policy.logits_t(env.obs).permute(2, 0, 1).float()
    assert (logits - eager).abs().max().item() < 0.05


@pytest.mark.parametrize("n_actions", [13, 7, 24])
def test_fused_policy_other_action_counts(cuda_lib, fused_impl, n_actions):
    """13 actions (SimplifyAction) has its own instantiation, other counts take the generic one."""
    from pikazoo_b200.policy import MLPPolicy, sample_reference

    n = 20_000
    env = _played_env(n)
    policy = MLPPolicy(n_actions=n_actions, device=env.device, seed=5)
    logits = torch.full((n, 2, n_actions), float("nan"), device="cuda")
    a = policy.act_fused(env.obs, step=3, seed=9, logits_out=logits)
    ref = _reference_logits(policy, env.obs)
    assert (logits - ref).abs().max().item() < 2e-3
    expect = _expected_samples(fused_impl, logits.cpu().numpy(), 9, 3, 0)
    assert int((a.cpu().numpy() != expect).sum()) <= 8 and int(a.max()) < n_actions
    g = policy.act_fused(env.obs, step=3, greedy=True, logits_out=logits)
    assert np.array_equal(g.cpu().numpy().astype(np.int64), sample_reference(logits.cpu().numpy(), None))


def test_fused_policy_sampling_is_the_documented_sampler(cuda_lib, fused_impl):
    """The sampled actions are a pure function of the logits and the counters, restated in numpy: tcgen05 —
    inversion of the cumulative distribution with one uniform per (env, agent); mma.sync — argmax(logits + Gumbel
    noise). The kernels take exp2 / log2 from the hardware approximations, so weights and keys differ from numpy's
    by a few float32 ulp: every mismatch must be a near-tie (a target within 1e-4 of a boundary of the cumulative
    distribution, two keys within 1e-4), and there must be next to none."""
    from pikazoo_b200.policy import MLPPolicy, _counter_uniform, gumbel_noise_reference

    n = 50_000
    env = _played_env(n)
    policy = MLPPolicy(device=env.device, seed=4)
    logits = torch.empty((n, 2, 18), device="cuda")
    for step, seed, first in ((0, 0, 0), (7, 123456789, 10**6), (2**40, 2**63 + 5, 3)):
        a = policy.act_fused(env.obs, step=step, seed=seed, first_env=first, logits_out=logits).cpu().numpy()
        lg = logits.cpu().numpy()
        expect = _expected_samples(fused_impl, lg, seed, step, first)
        bad = np.argwhere(a != expect)
        assert len(bad) <= 2 * n * 2e-4, len(bad)
        if fused_impl == "tcgen05":
            p = np.exp(lg.astype(np.float64) - lg.max(axis=-1, keepdims=True))
            cdf = np.cumsum(p, axis=-1) / p.sum(axis=-1, keepdims=True)
            u = _counter_uniform(seed, step, first, n, 1)[..., 0].astype(np.float64)
            for e, ag in bad:
                assert np.abs(cdf[e, ag] - u[e, ag]).min() < 1e-4 and abs(int(a[e, ag]) - int(expect[e, ag])) == 1
        else:
            keys = lg + gumbel_noise_reference(seed, step, first, n, 18)
            for e, ag in bad:
                top = np.sort(keys[e, ag])[-2:]
                assert top[1] - top[0] < 1e-4
    # a different step gives different samples; the same counters give the same ones
    a0 = policy.act_fused(env.obs, step=1).clone()
    assert not torch.equal(a0, policy.act_fused(env.obs, step=2))
    assert torch.equal(a0, policy.act_fused(env.obs, step=1))


def test_fused_policy_samples_follow_softmax(cuda_lib, fused_impl):
    """Statistics: after reset every env shows the same observation, so N envs are N independent samples of
    one categorical distribution; the frequencies must match softmax(logits) within five standard errors."""
    import pikazoo_b200
    from pikazoo_b200.policy import MLPPolicy

    n = 400_000
    env = pikazoo_b200.PikaVecEnv(n, seed=1, obs_dtype=torch.bfloat16, normalize_observation=True,
                                  obs_layout="feature_major", obs_feature_rows=40)
    env.reset()
    assert bool((env.obs == env.obs[:, :, :1]).all())
    policy = MLPPolicy(device=env.device, seed=9)
    with torch.no_grad():
        policy.w2.mul_(4.0)  # a peaked distribution
    logits = torch.empty((n, 2, 18), device="cuda")
    a = policy.act_fused(env.obs, step=11, seed=5, logits_out=logits)
    p = torch.softmax(logits[0].double(), dim=1).cpu().numpy()  # [2, 18]
    for agent in range(2):
        freq = np.bincount(a[:, agent].cpu().numpy(), minlength=18) / n
        se = np.sqrt(p[agent] * (1 - p[agent]) / n)
        assert np.all(np.abs(freq - p[agent]) < 5 * se + 1e-6), (agent, freq, p[agent])


def test_fused_actor_rollout_loop_matches_oracle(cuda_lib, fused_impl):
    """configs[4] with the fused actor: uint8 actions straight from the policy kernel into the step kernel; the
    oracle replays them."""
    import pikazoo_b200
    from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout

    n, steps = 8192, 400
    cfg = dict(winning_score=5, serve="random")
    env = pikazoo_b200.PikaVecEnv(n, seed=16, obs_dtype=torch.bfloat16, normalize_observation=True,
                                  action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40, **cfg)
    orc = po.OracleVecEnv(n, seed=16, **cfg)
    actor = FusedActor(MLPPolicy(device=env.device, seed=3), env, seed=77)
    env.reset(), orc.reset()
    seen = {"done": 0, "distinct": set()}

    def mirror(t, actions, obs, reward, done):
        a = actions.cpu().numpy()
        assert a.dtype == np.uint8 and a.max() < 18
        orc.step(a.astype(np.int32))
        assert np.array_equal(done.cpu().numpy(), orc.done.astype(bool)), t
        seen["done"] += int(done.sum())
        seen["distinct"] |= set(np.unique(a).tolist())

    policy_rollout(env, actor, steps, on_step=mirror)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    assert seen["done"] > n // 4 and len(seen["distinct"]) == 18


def test_fused_policy_rejects_bad_arguments(cuda_lib, fused_impl):
    from pikazoo_b200 import _lib
    from pikazoo_b200.policy import MLPPolicy

    policy = MLPPolicy(device="cuda", seed=1)
    with pytest.raises(ValueError):
        policy.act_fused(torch.zeros((64, 2, 35), dtype=torch.bfloat16, device="cuda"), step=0)  # env-major
    obs = torch.zeros((2, 40, 64), dtype=torch.bfloat16, device="cuda")
    out = torch.zeros((64, 2), dtype=torch.uint8, device="cuda")
    L = _lib.load()
    args = [obs.data_ptr(), 64, 64, 40, policy.w1.data_ptr(), 72, 40, policy.w2.data_ptr(), 18, 72, 0, 0, 0,
            out.data_ptr(), _lib.ACT_U8, 0, None, None]
    assert L.pz_policy_mlp_act(*args) == 0
    for pos, bad in ((6, 49), (5, 81), (8, 25), (9, 81), (3, 39), (14, 7), (2, 63), (0, None)):
        b = list(args)
        b[pos] = bad
        assert L.pz_policy_mlp_act(*b) < 0, pos


@pytest.mark.parametrize("n,n_actions", [(128, 18), (1000, 18), (4096 + 8, 13), (1 << 17, 18), (300_001, 7)])
def test_fused_policy_implementations_agree(cuda_lib, n, n_actions):
    """The tcgen05 kernel and the warp-level mma.sync kernel evaluate the same network: same logits (both accumulate
    the same bf16 products in fp32, sixteen features at a time) and therefore the same greedy actions, including
    ragged last tiles and batches smaller than one tile per SM. Their categorical samplers differ (inversion of the
    cumulative distribution / Gumbel arg-max); each returns what its numpy restatement returns."""
    from pikazoo_b200.policy import MLPPolicy

    env = _played_env(n)
    policy = MLPPolicy(n_actions=n_actions, device=env.device, seed=8)
    with torch.no_grad():  # non-zero biases
        policy.w1[:, : policy.hidden, policy.ONES_ROW] = 0.25
        policy.w2[:, :, policy.hidden] = -0.5
    got = {}
    prev = cuda_lib.pz_policy_select(0)
    try:
        for name, code in IMPLS.items():
            assert cuda_lib.pz_policy_select(code) in (0, 1)
            logits = torch.full((n, 2, n_actions), float("nan"), device="cuda")
            sampled = policy.act_fused(env.obs, step=6, seed=2, first_env=5, logits_out=logits).clone()
            greedy = policy.act_fused(env.obs, step=6, greedy=True, action_dtype=torch.int64).clone()
            expect = _expected_samples(name, logits.cpu().numpy(), 2, 6, 5)
            assert float((sampled.cpu().numpy() != expect).mean()) <= 2e-4, name
            got[name] = (logits, greedy)
    finally:
        cuda_lib.pz_policy_select(prev)
    (l0, g0), (l1, g1) = got["tcgen05"], got["mma_sync"]
    assert not torch.isnan(l0).any() and not torch.isnan(l1).any()
    assert float((l0 - l1).abs().max()) <= 1e-5  # measured: bit-identical
    assert float((g0 != g1).float().mean()) <= 1e-4
    assert cuda_lib.pz_policy_select(7) == -1  # unknown code: refused, selection unchanged
