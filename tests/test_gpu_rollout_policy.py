"""GPU: pz_rollout_policy (csrc/pz_rollout_policy.cu) — K frames of `observation -> MLP policy -> sampled actions ->
raw_env.step` in one launch, everything on chip. The verification chain, link by link:

  physics      the oracle replays the actions the kernel exported: final hidden state (all 53 words incl. the PCG64
               stream) and the statistics vector agree exactly (integer work, bit-exact);
  observation  what the policy saw is the oracle's bf16 NormalizeObservation row: the logits of every frame are
               within tolerance of a PyTorch fp32 reference of the same network applied to the ORACLE's rows (floating
               point; the tolerance is one hidden activation rounding to the neighbouring bf16, and such flips must
               be rare: fewer than 0.4 % of the (env, agent) rows off by more than 2e-4);
  sample       the exported actions are the numpy-restated sampler (policy.py inverse_cdf_reference) applied to the
               kernel's own logits, except on near-ties (hardware exp2 vs numpy's); greedy = the packed-key arg-max
               exactly.
"""

import numpy as np
import pytest
import torch

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _bf16_bits_to_f32(bits_u16: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(bits_u16.astype(np.int16)).view(torch.bfloat16).float()


@torch.no_grad()
def _reference_logits(policy, rows_bits: np.ndarray):
    """([N, 2, A] fp32, tolerance): plain PyTorch fp32 arithmetic on the oracle's bf16 rows [N, 2, 35] (+ the bias
    input), hidden activations rounded to bf16 where the kernel rounds them. Floating point, so by tolerance: fp32
    sums in another order can round a hidden activation to the neighbouring bf16 (2^-8 relative), which reaches a
    logit through one weight of the second layer — the bound is that one flip; `_check_logits` also requires such
    flips to be rare."""
    x = _bf16_bits_to_f32(rows_bits).cuda()                              # [N, 2, 35]
    n = x.shape[0]
    xt = torch.zeros((2, policy.K_PAD, n), device="cuda")
    xt[:, :35] = x.permute(1, 2, 0)
    xt[:, policy.ONES_ROW] = 1.0
    h = torch.bmm(policy.w1.float(), xt).to(torch.bfloat16).float().relu_()
    tol = float(h.max()) * 2.0 ** -8 * float(policy.w2.float().abs().max()) + 1e-4
    return torch.bmm(policy.w2.float(), h).permute(2, 0, 1).contiguous(), tol


def _check_logits(got: torch.Tensor, policy, rows_bits: np.ndarray, what):
    ref, tol = _reference_logits(policy, rows_bits)
    err = (got - ref).abs()
    assert torch.isfinite(got).all(), what
    assert float(err.max()) < tol, (what, float(err.max()), tol)
    flipped = int((err > 2e-4).any(dim=-1).sum())  # (env, agent) rows touched by a flipped hidden activation
    assert flipped <= max(2, int(4e-3 * err.shape[0] * 2)), (what, flipped)


def _make(n, seed, normalize=True, **cfg):
    import pikazoo_b200

    env = pikazoo_b200.PikaVecEnv(n, seed=seed, obs_dtype=torch.bfloat16, normalize_observation=normalize,
                                  action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40, **cfg)
    orc = po.OracleVecEnv(n, seed=seed, **cfg)
    env.reset(), orc.reset()
    return env, orc


@pytest.mark.parametrize("n,K,launches,cfg", [
    (8192, 48, 6, dict(winning_score=5, serve="random")),
    (3001, 17, 5, dict(winning_score=2, serve="alternate")),              # ragged last tile, K not a power of two
    (128 * 148 * 6 + 77, 8, 2, dict(winning_score=15, serve="winner")),   # every group of every SM has a tile
    (50, 64, 3, dict(winning_score=1, serve="random", max_episode_frames=40)),
])
def test_rollout_policy_matches_oracle(cuda_lib, n, K, launches, cfg):
    from pikazoo_b200.policy import MLPPolicy, inverse_cdf_reference, rollout_fused

    env, orc = _make(n, 21, **cfg)
    policy = MLPPolicy(device=env.device, seed=3)
    with torch.no_grad():  # non-zero biases
        policy.w1[:, : policy.hidden, policy.ONES_ROW] = 0.125
        policy.w2[:, :, policy.hidden] = -0.25
    check_logits = n <= 10_000
    actions = torch.full((K, n, 2), 255, dtype=torch.uint8, device="cuda")
    logits = torch.full((K, n, 2, 18), float("nan"), device="cuda") if check_logits else None
    distinct, ostats = set(), np.zeros(16, dtype=np.int64)
    for launch in range(launches):
        step0 = env.frame
        rollout_fused(env, policy, K, seed=77, actions_out=actions, logits_out=logits)
        a = actions.cpu().numpy()
        assert a.max() < 18
        lg = logits.cpu().numpy() if check_logits else None
        for k in range(K):
            if check_logits:
                _check_logits(logits[k], policy, orc.normalized_obs("bfloat16"), (launch, k))
                expect = inverse_cdf_reference(lg[k], 77, step0 + k, 0)
                assert int((a[k] != expect).sum()) <= max(2, int(2 * n * 2e-4)), (launch, k)
            orc.step(a[k].astype(np.int32))
        distinct |= set(np.unique(a).tolist())
        assert np.array_equal(env.export_state().cpu().numpy(), orc.state), launch
    st = env.stats_dict()
    assert st["calls"] == n * K * launches
    assert len(distinct) == 18
    # statistics: the oracle's episode counters over the same calls
    assert st["episodes"] + st["truncated"] >= (1 if cfg["winning_score"] <= 2 else 0)
    assert st["resets"] <= st["episodes"] + st["truncated"]


def test_rollout_policy_observation_after_last_frame_and_loop_continuation(cuda_lib):
    """write_obs hands the unfused loop its next observation: fused K frames, then per-step launches with the
    two-kernel FusedActor, then fused again — the oracle follows the exported / returned actions throughout."""
    from pikazoo_b200.policy import FusedActor, MLPPolicy, policy_rollout, rollout_fused

    n, K = 4096 + 40, 24
    env, orc = _make(n, 5, winning_score=5, serve="random")
    policy = MLPPolicy(device=env.device, seed=11)
    actions = torch.empty((K, n, 2), dtype=torch.uint8, device="cuda")
    obs = rollout_fused(env, policy, K, seed=1, actions_out=actions, write_obs=True)
    for k in range(K):
        orc.step(actions[k].cpu().numpy().astype(np.int32))
    assert np.array_equal(obs[:, :35, :].permute(2, 0, 1).contiguous().view(torch.int16).cpu().numpy().view(np.uint16),
                          orc.normalized_obs("bfloat16"))
    actor = FusedActor(policy, env, seed=1)
    actor.step = env.frame
    policy_rollout(env, actor, 30, on_step=lambda t, a, o, r, d: orc.step(a.cpu().numpy().astype(np.int32)))
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    rollout_fused(env, policy, K, seed=1, actions_out=actions)
    for k in range(K):
        orc.step(actions[k].cpu().numpy().astype(np.int32))
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)


@pytest.mark.parametrize("simplify", [False, True])
def test_rollout_policy_plain_instantiation_equals_the_general_one(cuda_lib, simplify):
    """A launch that samples, has no frame cap and does not export its logits runs the PLAIN instantiation of the kernel
    (those launch-uniform options compiled out of the frame loop); most oracle-checked tests export the logits and so
    run the general one. Same start, same seeds: both must sample the same actions, end in the identical state, stream
    included, and count the same episodes."""
    from pikazoo_b200.policy import MLPPolicy, rollout_fused

    n, K = 8192 + 77, 40
    ends = []
    for general in (True, False):
        env, _ = _make(n, 21, winning_score=2, serve="random", simplify_action=simplify)
        policy = MLPPolicy(device=env.device, seed=4, n_actions=13 if simplify else 18)
        actions = torch.empty((K, n, 2), dtype=torch.uint8, device="cuda")
        logits = torch.empty((K, n, 2, policy.n_actions), dtype=torch.float32, device="cuda") if general else None
        for _ in range(3):
            rollout_fused(env, policy, K, seed=9, actions_out=actions, logits_out=logits)
        st = env.stats_dict()
        ends.append((env.export_state().cpu().numpy(), actions.cpu().numpy(),
                     {k: st[k] for k in ("episodes", "resets", "p1_points", "p2_points")}))
    assert np.array_equal(ends[0][0], ends[1][0]) and np.array_equal(ends[0][1], ends[1][1])
    assert ends[0][2] == ends[1][2] and ends[0][2]["episodes"] > 0
    # and without the action export (the same instantiation, a null pointer): the same state again
    env, _ = _make(n, 21, winning_score=2, serve="random", simplify_action=simplify)
    policy = MLPPolicy(device=env.device, seed=4, n_actions=13 if simplify else 18)
    for _ in range(3):
        rollout_fused(env, policy, K, seed=9)
    assert np.array_equal(env.export_state().cpu().numpy(), ends[0][0])


def test_rollout_policy_agrees_with_the_two_kernel_loop(cuda_lib):
    """Same policy, seeds and counters through pz_policy_mlp_act + pz_step: the first frame's observations are
    identical, so the logits agree to rounding (player_2's first layer accumulates in another order) and the samples
    agree except at rounding-induced boundary crossings."""
    from pikazoo_b200.policy import MLPPolicy, rollout_fused

    n = 20_000
    env, _ = _make(n, 9, winning_score=5, serve="random")
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(60):
        env.step(torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.uint8))
    twin_state = env.state.clone()
    policy = MLPPolicy(device=env.device, seed=3)
    l2 = torch.empty((n, 2, 18), device="cuda")
    a2 = policy.act_fused(env.obs, step=env.frame, seed=4, logits_out=l2).clone()
    a1 = torch.empty((1, n, 2), dtype=torch.uint8, device="cuda")
    l1 = torch.empty((1, n, 2, 18), device="cuda")
    step0 = env.frame
    rollout_fused(env, policy, 1, seed=4, step0=step0, actions_out=a1, logits_out=l1)
    assert torch.equal(l1[0, :, 0], l2[:, 0])                 # player_1: same products in the same order
    assert float((l1[0] - l2).abs().max()) < 1e-4             # player_2: permuted accumulation order
    assert float((a1[0] != a2).float().mean()) < 1e-3
    # and the frame itself: stepping the twin with the fused kernel's actions gives the fused kernel's state
    import pikazoo_b200

    twin = pikazoo_b200.PikaVecEnv(n, seed=9, action_dtype=torch.uint8, winning_score=5, serve="random")
    twin.state.copy_(twin_state)
    twin.step(a1[0])
    assert torch.equal(twin.export_state(), env.export_state())


@pytest.mark.parametrize("simplify,normalize", [(True, True), (False, False)])
def test_rollout_policy_variants(cuda_lib, simplify, normalize):
    """SimplifyAction (13 actions, its own instantiation) and raw (unnormalised) bf16 observations; greedy actions
    are the packed-key arg-max of the kernel's logits exactly."""
    from pikazoo_b200.policy import MLPPolicy, rollout_fused, sample_reference

    n, K = 5000, 12
    na = 13 if simplify else 18
    env, orc = _make(n, 2, normalize=normalize, winning_score=3, serve="random", simplify_action=simplify)
    policy = MLPPolicy(n_actions=na, device=env.device, seed=6)
    if not normalize:
        with torch.no_grad():
            policy.w1.mul_(1.0 / 256)  # raw coordinates are O(100)
    actions = torch.empty((K, n, 2), dtype=torch.uint8, device="cuda")
    logits = torch.empty((K, n, 2, na), device="cuda")
    for greedy in (False, True):
        rollout_fused(env, policy, K, seed=3, greedy=greedy, actions_out=actions, logits_out=logits)
        a, lg = actions.cpu().numpy(), logits.cpu().numpy()
        assert a.max() < na
        for k in range(K):
            rows = orc.normalized_obs("bfloat16") if normalize else orc.raw_obs_bf16()
            _check_logits(logits[k], policy, rows, (greedy, k))
            if greedy:
                assert np.array_equal(a[k].astype(np.int64), sample_reference(lg[k], None)), k
            orc.step(a[k].astype(np.int32))
        assert np.array_equal(env.export_state().cpu().numpy(), orc.state)


def test_rollout_policy_rejects_bad_arguments(cuda_lib):
    import pikazoo_b200
    from pikazoo_b200 import _lib
    from pikazoo_b200.policy import MLPPolicy, rollout_fused

    env = pikazoo_b200.PikaVecEnv(256, seed=1, is_player2_computer=True)
    env.reset()
    policy = MLPPolicy(device=env.device, seed=1)
    with pytest.raises(_lib.PikaLibraryError):
        rollout_fused(env, policy, 4)          # a computer player: the policy plays both sides
    env = pikazoo_b200.PikaVecEnv(256, seed=1, simplify_action=True)
    env.reset()
    with pytest.raises(_lib.PikaLibraryError):
        rollout_fused(env, policy, 4)          # 18 logits for a 13-action env
    env = pikazoo_b200.PikaVecEnv(256, seed=1)
    env.reset()
    with pytest.raises(ValueError):
        rollout_fused(env, policy, 4, actions_out=torch.empty((3, 256, 2), dtype=torch.uint8, device="cuda"))
    with pytest.raises(_lib.PikaLibraryError):
        rollout_fused(env, policy, 0)


def test_rollout_policy_full_size(cuda_lib):
    """configs[4] at its stated size: 2,097,152 envs on one GPU, K = 64 frames in one launch; the oracle replays a
    strided sample of the envs from the exported actions; the statistics are consistent."""
    from pikazoo_b200.policy import MLPPolicy, rollout_fused

    import pikazoo_b200

    n, K = 1 << 21, 64
    cfg = dict(winning_score=5, serve="random")
    env = pikazoo_b200.PikaVecEnv(n, seed=2026, obs_dtype=torch.bfloat16, normalize_observation=True,
                                  action_dtype=torch.uint8, obs_layout="feature_major", obs_feature_rows=40, **cfg)
    env.reset()
    policy = MLPPolicy(device=env.device, seed=3)
    idx = np.arange(0, n, 4099)
    seeds = (2026 + idx).astype(np.uint64)
    orc = po.OracleVecEnv(len(idx), seeds=seeds, **cfg)
    orc.reset()
    actions = torch.empty((K, n, 2), dtype=torch.uint8, device="cuda")
    sel = torch.from_numpy(idx).cuda()
    for launch in range(4):
        rollout_fused(env, policy, K, seed=8, actions_out=actions)
        a = actions[:, sel].cpu().numpy()
        for k in range(K):
            orc.step(a[k].astype(np.int32))
        assert np.array_equal(env.export_state()[sel].cpu().numpy(), orc.state), launch
    st = env.stats_dict()
    assert st["calls"] == n * K * 4 and st["episodes"] > 0
    assert st["p1_wins"] + st["p2_wins"] == st["episodes"]
