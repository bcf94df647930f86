"""GPU: differential fuzz over the CONFIGURATION space. The per-feature tests pin every option on its own; this one
draws random combinations — computer-player mask x serve mode x winning score x fused wrappers x observation dtype x
observation layout x action dtype x reward dtype x truncation x auto-reset x ragged batch sizes x landing tables —
and runs each in lock-step with the oracle (per-step kernel, then a K-frame rollout): observations as bit patterns,
rewards, flags, episode statistics and the full hidden state must agree on every compared step."""

import numpy as np
import pytest
import torch

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy

pytestmark = pytest.mark.gpu

_OBS = [(torch.int32, np.int32, False), (torch.int16, np.int16, False), (torch.float32, np.float32, False),
        (torch.float32, np.float32, True), (torch.float16, np.float16, True), (torch.bfloat16, "bfloat16", True),
        (torch.bfloat16, "bfloat16", False), (torch.float64, np.float64, True)]


def _bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _draw_case(rng):
    shaped = (tuple(float(x) for x in rng.choice([0.0, 0.125, -0.25, 0.1, -0.3, 1.0], 8)), int(rng.integers(100, 330)),
              int(rng.integers(60, 240)))
    cfg = dict(
        winning_score=int(rng.choice([1, 2, 3, 5, 15])),
        serve=str(rng.choice(["winner", "alternate", "random"])),
        is_player1_computer=bool(rng.integers(0, 2)), is_player2_computer=bool(rng.integers(0, 2)),
        simplify_action=bool(rng.integers(0, 2)),
        reward_by_ball_position=shaped if rng.integers(0, 2) else None,
        reward_in_normal_state=float(rng.choice([-0.01, 0.5])) if rng.integers(0, 3) == 0 else None,
        normal_state_first=bool(rng.integers(0, 2)),
        max_episode_frames=int(rng.choice([0, 0, 37, 150])),
    )
    obs_t, obs_np, normalize = _OBS[int(rng.integers(0, len(_OBS)))]
    extra = dict(
        obs_dtype=obs_t, normalize_observation=normalize,
        obs_layout=str(rng.choice(["env_major", "feature_major"])),
        action_dtype=[torch.int32, torch.int64, torch.uint8][int(rng.integers(0, 3))],
        reward_dtype=[torch.float32, torch.float64][int(rng.integers(0, 2))],
        autoreset=bool(rng.integers(0, 4) != 0), record_episode_statistics=True,
        landing_tables=bool(rng.integers(0, 2)),
    )
    if extra["obs_layout"] == "feature_major":
        extra["obs_feature_rows"] = int(rng.choice([35, 40]))
    n = int(rng.choice([1, 31, 32, 33, 127, 640, 1000, 4097]))
    return n, cfg, extra, obs_np, normalize


@pytest.mark.parametrize("case", range(96))
def test_random_configuration_matches_oracle(cuda_lib, case):
    import pikazoo_b200

    rng = np.random.default_rng(1000 + case)
    n, cfg, extra, obs_np, normalize = _draw_case(rng)
    seed = int(rng.integers(0, 2**40))
    env = pikazoo_b200.PikaVecEnv(n, seed=seed, **cfg, **extra)
    orc = po.OracleVecEnv(n, seed=seed, autoreset=extra["autoreset"], **cfg)
    n_act = 13 if cfg["simplify_action"] else 18
    label = (case, n, cfg, {k: str(v) for k, v in extra.items()})

    def check_obs(t):
        got = env.obs
        if extra["obs_layout"] == "feature_major":
            assert not bool(got[:, 35:, :].any()), label  # padding rows are never written
            got = got[:, :35, :].permute(2, 0, 1).contiguous()
        want = po.convert_obs(orc.obs, obs_np, normalize)
        assert np.array_equal(_bits(got), want), (label, t)

    env.reset(), orc.reset()
    check_obs(-1)
    steps = 260
    for t in range(steps):
        a = synth_actions_numpy(seed & 0xFFFF, 0, n, t, n_act)
        obs, rew, done = env.step(torch.from_numpy(a).to("cuda", dtype=extra["action_dtype"]))
        orc.step(a)
        assert np.array_equal(done.cpu().numpy(), orc.done.astype(bool)), (label, t)
        if t % 20 == 0 or t == steps - 1:
            check_obs(t)
            want_rew = orc.reward if extra["reward_dtype"] == torch.float64 else orc.reward.astype(np.float32)
            assert np.array_equal(rew.cpu().numpy(), want_rew), (label, t)
            assert np.array_equal(env.episode_return.cpu().numpy(), orc.episode_return), (label, t)
            assert np.array_equal(env.episode_length.cpu().numpy(), orc.episode_length), (label, t)
            if cfg["max_episode_frames"]:
                assert np.array_equal(env.truncated.cpu().numpy(), orc.truncated.astype(bool)), (label, t)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state), label
    if extra["autoreset"]:  # the K-frame rollout always auto-resets
        both_ai = cfg["is_player1_computer"] and cfg["is_player2_computer"]
        mode, name = (0, "noop") if both_ai else (1, "synth")
        env.rollout(48, actions=name, action_seed=99, write_obs=True)
        orc.rollout(48, action_mode=mode, action_seed=99, first_env=0, frame0=steps)
        assert np.array_equal(env.export_state().cpu().numpy(), orc.state), label
        orc.current_obs()
        check_obs("rollout")


_PLAIN_OBS = [(torch.int32, np.int32, False), (torch.int16, np.int16, False), (torch.float32, np.float32, True),
              (torch.float16, np.float16, True), (torch.bfloat16, "bfloat16", True), (torch.float64, np.float64, True)]


@pytest.mark.parametrize("case", range(40))
def test_random_plain_configuration_matches_oracle(cuda_lib, case):
    """The same fuzz restricted to what the PLAIN instantiations of the per-step kernels serve (no wrappers, no frame
    cap, no episode statistics, auto-reset, float32 rewards, normalised rows iff floating point): every case above asks
    for episode statistics and therefore runs the general kernels; these run the PLAIN ones."""
    import pikazoo_b200

    rng = np.random.default_rng(5000 + case)
    cfg = dict(winning_score=int(rng.choice([1, 2, 3, 15])), serve=str(rng.choice(["winner", "alternate", "random"])),
               is_player1_computer=bool(rng.integers(0, 2)), is_player2_computer=bool(rng.integers(0, 2)))
    obs_t, obs_np, normalize = _PLAIN_OBS[int(rng.integers(0, len(_PLAIN_OBS)))]
    layout = str(rng.choice(["env_major", "feature_major"]))
    extra = dict(obs_dtype=obs_t, normalize_observation=normalize, obs_layout=layout,
                 action_dtype=[torch.int32, torch.int64, torch.uint8][int(rng.integers(0, 3))],
                 landing_tables=bool(rng.integers(0, 2)))
    if layout == "feature_major":
        extra["obs_feature_rows"] = int(rng.choice([35, 40]))
    n = int(rng.choice([1, 31, 33, 128, 640, 1000, 4097]))
    seed = int(rng.integers(0, 2**40))
    env = pikazoo_b200.PikaVecEnv(n, seed=seed, **cfg, **extra)
    orc = po.OracleVecEnv(n, seed=seed, **cfg)
    label = (case, n, cfg, {k: str(v) for k, v in extra.items()})

    def check_obs(t):
        got = env.obs
        if layout == "feature_major":
            got = got[:, :35, :].permute(2, 0, 1).contiguous()
        assert np.array_equal(_bits(got), po.convert_obs(orc.obs, obs_np, normalize)), (label, t)

    env.reset(), orc.reset()
    check_obs(-1)
    steps = 300
    for t in range(steps):
        a = synth_actions_numpy(seed & 0xFFFF, 0, n, t, 18)
        obs, rew, done = env.step(torch.from_numpy(a).to("cuda", dtype=extra["action_dtype"]))
        orc.step(a)
        assert np.array_equal(done.cpu().numpy(), orc.done.astype(bool)), (label, t)
        if t % 25 == 0 or t == steps - 1:
            check_obs(t)
            assert np.array_equal(rew.cpu().numpy(), orc.reward.astype(np.float32)), (label, t)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state), label
    st = env.stats_dict()
    assert st["bad_actions"] == 0 and st["frozen"] == 0 and st["truncated"] == 0
