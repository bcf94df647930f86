"""GPU: the fused forms of the reference's remaining wrappers through the
C ABI — NormalizeObservation and the observation dtypes, RewardInNormalState inside / outside
RewardByBallPosition, RecordEpisodeStatistics, truncation — against the golden sessions recorded
from the reference's own wrapper classes and against the oracle in lock-step. Bit-exact: float64
observations and rewards are compared as bit patterns; float32 = float32(float64 value);
float16 / bfloat16 = round-to-nearest-even of that float32."""

import numpy as np
import pytest
import torch

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy
from tests.helpers import replay_group
from tests.test_oracle_golden import WRAPPER_GROUPS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pz(cuda_lib):
    import pikazoo_b200

    return pikazoo_b200


class CudaWrapperStepper:
    """PikaVecEnv configured from a golden group's wrapper stack."""

    def __init__(self, n, base_seed, cfg):
        import pikazoo_b200

        cfg = dict(cfg)
        self.normalized = bool(cfg.get("normalize_observation"))
        self.env = pikazoo_b200.PikaVecEnv(
            n, device="cuda", seed=base_seed, autoreset=True, reward_dtype=torch.float64,
            obs_dtype=torch.float64 if self.normalized else torch.int32, **cfg)

    def reset(self):
        return self.env.reset().cpu().numpy()

    def step(self, actions):
        obs, rew, done = self.env.step(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32)).cuda())
        return obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()

    def scores(self):
        return self.env.scores().cpu().numpy()

    def final_state(self):
        return self.env.export_state().cpu().numpy()

    def episode(self):
        return self.env.episode_return.cpu().numpy(), self.env.episode_length.cpu().numpy()


@pytest.mark.parametrize("name", WRAPPER_GROUPS)
def test_cuda_replays_reference_wrapper_stacks(pz, golden_wrappers, name):
    group = next(g for g in golden_wrappers["groups"] if g["name"] == name)
    bad = replay_group(group, CudaWrapperStepper)
    assert not bad, bad[:5]


TORCH_TO_NP = {torch.int32: np.int32, torch.int16: np.int16, torch.float32: np.float32, torch.float16: np.float16,
               torch.float64: np.float64, torch.bfloat16: "bfloat16"}


def _bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    a = t.cpu().numpy()
    return a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize]) if a.dtype.kind == "f" else a


def _expected_bits(obs_i32, np_dtype, normalize):
    e = po.convert_obs(obs_i32, np_dtype, normalize)
    if isinstance(np_dtype, str):
        return e
    return e.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[e.dtype.itemsize]) if e.dtype.kind == "f" else e


@pytest.mark.parametrize("dtype,normalize", [
    (torch.int16, False), (torch.float32, False), (torch.float32, True), (torch.float16, True),
    (torch.bfloat16, True), (torch.float64, True), (torch.float64, False), (torch.float16, False),
])
@pytest.mark.parametrize("n", [4096, 77])  # full warps through the bulk copy, and a ragged tail
def test_observation_dtypes(pz, dtype, normalize, n):
    cfg = dict(winning_score=3, serve="random", is_player2_computer=True)
    env = pz.PikaVecEnv(n, seed=123, obs_dtype=dtype, normalize_observation=normalize, **cfg)
    orc = po.OracleVecEnv(n, seed=123, **cfg)
    np_dtype = TORCH_TO_NP[dtype]
    assert np.array_equal(_bits(env.reset()), _expected_bits(orc.reset(), np_dtype, normalize))
    for t in range(400):
        a = synth_actions_numpy(9, 0, n, t, 18)
        obs, _, _ = env.step(torch.from_numpy(a).cuda())
        orc.step(a)
        if t % 7 == 0 or t == 399:
            assert obs.dtype == dtype
            assert np.array_equal(_bits(obs), _expected_bits(orc.obs, np_dtype, normalize)), t
    # the rollout kernel's final observation goes through the same conversion
    obs = env.rollout(16, actions="synth", action_seed=5, write_obs=True)
    orc.rollout(16, action_mode=1, action_seed=5, first_env=0, frame0=400)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    assert np.array_equal(_bits(obs), _expected_bits(orc.current_obs(), np_dtype, normalize))


@pytest.mark.parametrize("dtype,normalize", [(torch.int32, False), (torch.int16, False), (torch.float32, True),
                                             (torch.bfloat16, True), (torch.float16, False), (torch.float64, True)])
@pytest.mark.parametrize("n,cfg", [(4096, dict(winning_score=3, serve="random", is_player2_computer=True)),
                                   (1000 + 7, dict(winning_score=2))])
def test_feature_major_layout(pz, dtype, normalize, n, cfg):
    """obs [2, rows, N]: the same values as the env-major rows, transposed; padding rows stay zero."""
    rows = 40
    env = pz.PikaVecEnv(n, seed=321, obs_dtype=dtype, normalize_observation=normalize, obs_layout="feature_major",
                        obs_feature_rows=rows, **cfg)
    orc = po.OracleVecEnv(n, seed=321, **cfg)
    np_dtype = TORCH_TO_NP[dtype]

    def check(obs, where):
        assert tuple(obs.shape) == (2, rows, n) and obs.dtype == dtype
        exp = _expected_bits(orc.current_obs(), np_dtype, normalize)          # [n, 2, 35]
        got = _bits(obs)                                                      # [2, rows, n]
        assert np.array_equal(got[:, :35, :], np.transpose(exp, (1, 2, 0))), where
        assert not got[:, 35:, :].any(), where

    orc.reset()
    check(env.reset(), "reset")
    for t in range(300):
        a = synth_actions_numpy(9, 0, n, t, 18)
        obs, _, _ = env.step(torch.from_numpy(a).cuda())
        orc.step(a)
        if t % 13 == 0 or t == 299:
            check(obs, t)
    obs = env.rollout(16, actions="synth", action_seed=5, write_obs=True)
    orc.rollout(16, action_mode=1, action_seed=5, first_env=0, frame0=300)
    check(obs, "rollout")
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)


def test_normalize_needs_float_dtype(pz):
    with pytest.raises(TypeError):
        pz.PikaVecEnv(8, normalize_observation=True)


@pytest.mark.parametrize("first", [False, True])
def test_reward_in_normal_state_and_episode_statistics_lockstep(pz, first):
    n, steps = 4096, 1200
    cfg = dict(winning_score=2, serve="alternate", simplify_action=True,
               reward_by_ball_position=((0.5, 0, -0.5, 0.25, 0, 0.125, 0, -1), 216, 176),
               reward_in_normal_state=0.03125, normal_state_first=first)
    env = pz.PikaVecEnv(n, seed=77, reward_dtype=torch.float64, record_episode_statistics=True, **cfg)
    orc = po.OracleVecEnv(n, seed=77, **cfg)
    env.reset(), orc.reset()
    finished = 0
    for t in range(steps):
        a = synth_actions_numpy(3, 0, n, t, 13)
        _, rew, done = env.step(torch.from_numpy(a).cuda())
        orc.step(a)
        assert np.array_equal(rew.cpu().numpy(), orc.reward), t
        d = done.cpu().numpy()
        assert np.array_equal(d, orc.done.astype(bool))
        if d.any() or t % 100 == 0:
            assert np.array_equal(env.episode_return.cpu().numpy().view(np.uint64), orc.episode_return.view(np.uint64))
            assert np.array_equal(env.episode_length.cpu().numpy(), orc.episode_length)
            finished += int(d.sum())
    assert finished > n  # every env finished at least one game on average


def test_float32_reward_is_float32_of_the_double(pz):
    n = 2048
    cfg = dict(winning_score=2, reward_by_ball_position=((0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4), 216, 176),
               reward_in_normal_state=-0.001)
    env = pz.PikaVecEnv(n, seed=5, **cfg)  # default float32 rewards
    orc = po.OracleVecEnv(n, seed=5, **cfg)
    env.reset(), orc.reset()
    for t in range(300):
        a = synth_actions_numpy(8, 0, n, t, 18)
        _, rew, _ = env.step(torch.from_numpy(a).cuda())
        orc.step(a)
        assert np.array_equal(rew.cpu().numpy().view(np.uint32), orc.reward.astype(np.float32).view(np.uint32))


@pytest.mark.parametrize("autoreset", [True, False])
def test_truncation(pz, autoreset):
    """AI-vs-AI rallies can go on forever (seed 2 of config 1 never terminates): max_episode_frames."""
    n, cap = 512, 150
    cfg = dict(winning_score=15, is_player1_computer=True, is_player2_computer=True, max_episode_frames=cap)
    env = pz.PikaVecEnv(n, seed=0, autoreset=autoreset, record_episode_statistics=True, **cfg)
    orc = po.OracleVecEnv(n, seed=0, autoreset=autoreset, **cfg)
    env.reset(), orc.reset()
    for t in range(2 * cap + 10):
        obs, rew, done = env.step(None)
        orc.step(None)
        assert np.array_equal(obs.cpu().numpy(), orc.obs)
        assert np.array_equal(env.truncated.cpu().numpy(), orc.truncated.astype(bool)), t
        assert np.array_equal(env.episode_length.cpu().numpy(), orc.episode_length)
        assert not done.any()
        if t == cap - 1:
            assert env.truncated.all() and (env.episode_length == cap).all()
        if t == cap:
            assert bool(env.truncated.any()) == (not autoreset)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    s = env.stats_dict()
    assert s["truncated"] == (2 * n if autoreset else n) and s["episodes"] == 0
    assert s["frozen"] == (0 if autoreset else n * (cap + 10))


def test_truncation_in_rollout(pz):
    n, cap = 1000, 100
    cfg = dict(winning_score=15, is_player1_computer=True, is_player2_computer=True, max_episode_frames=cap)
    env = pz.PikaVecEnv(n, seed=3, **cfg)
    orc = po.OracleVecEnv(n, seed=3, **cfg)
    env.reset(), orc.reset()
    stats = np.zeros(9, dtype=np.int64)
    for _ in range(5):
        env.rollout(64)
        orc.rollout(64, stats=stats)
        assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    assert env.stats_dict()["truncated"] == stats[8] == n * 3  # 320 calls = 3 capped episodes + 3 resets + 17
    assert env.stats_dict()["resets"] == stats[7]


def test_facade_wrapper_stack_matches_reference_protocol(pz):
    """pikazoo_v0.env under the reference's wrapper classes' signatures, values from the golden file's
    first session (normalize + RewardInNormalState outside RewardByBallPosition + SimplifyAction + record)."""
    from oracle.make_golden import W1
    from pikazoo_b200 import wrappers as W

    seed = 7000
    env = pz.pikazoo_v0.env(winning_score=W1["winning_score"], serve=W1["serve"], seed=seed)
    env = W.SimplifyAction(env)
    add, xl, yl = W1["reward_by_ball_position"]
    env = W.RewardByBallPosition(env, add, xl, yl)
    env = W.RewardInNormalState(env, W1["reward_in_normal_state"])
    env = W.NormalizeObservation(env)
    env = W.RecordEpisodeStatistics(env)
    orc = po.OracleVecEnv(1, seed=seed, autoreset=False, **W1)
    obs, infos = env.reset()
    orc.reset()
    assert obs["player_1"].dtype == np.float64
    assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), orc.normalized_obs()[0])
    assert env.observation_space("player_1").dtype == np.float32 and env.action_space("player_2").n == 13
    t = 0
    while env.agents:
        a = synth_actions_numpy(0x5EED, 0, 1, t, 13)
        obs, rew, term, trunc, infos = env.step({"player_1": int(a[0, 0]), "player_2": int(a[0, 1])})
        orc.step(a)
        assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), orc.normalized_obs()[0])
        assert [rew["player_1"], rew["player_2"]] == orc.reward[0].tolist()
        t += 1
    assert infos["player_1"]["episode"] == {"r": orc.episode_return[0, 0], "l": t}
    assert infos["player_2"]["episode"]["r"] == orc.episode_return[0, 1]


def test_convert_single_agent(pz):
    from pikazoo_b200 import wrappers as W

    env = W.ConvertSingleAgent(pz.pikazoo_v0.env(winning_score=1, seed=4), "player_2")
    obs, info = env.reset()
    assert obs.shape == (35,) and info == {"score": [0, 0]}
    total, steps, term = 0, 0, False
    while not term and steps < 5000:
        obs, r, term, trunc, info = env.step(steps % 18)
        assert obs.shape == (35,) and trunc is False and isinstance(r, int)
        total += r
        steps += 1
    assert term and total in (-1, 1) and sum(info["score"]) == 1


def _orders_golden():
    import json
    import os

    from tests.conftest import ROOT

    with open(os.path.join(ROOT, "tests", "golden", "wrapper_orders.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["record_inside_rins", "record_inside_rbbp", "normalize_inside_rbbp", "rbbp_twice",
                                  "rins_twice", "normalize_twice", "mixed", "canonical"])
def test_facade_wrapper_orders_match_reference(pz, name):
    """Wrapper stacks in orders the fused kernel options do not cover (RecordEpisodeStatistics below a reward
    wrapper records unshaped rewards; RewardByBallPosition over NormalizeObservation reads normalised coordinates;
    two RewardByBallPosition add up; ...): tests/golden/wrapper_orders.json holds what the reference's own classes
    return for each stack (oracle/make_golden.py --orders); the facade, which fuses what the kernel reproduces at
    that position and runs the rest on the host, must return the same observations, rewards, terminations and
    episode statistics, hashed alike."""
    from oracle import wrapper_orders as wo
    from pikazoo_b200 import wrappers as W

    g = _orders_golden()
    assert g["env"] == wo.ENV_KW and g["action_seed"] == wo.ACTION_SEED
    stack = wo.STACKS[name]
    for sess in g["stacks"][name]["sessions"]:
        seed = sess["seed"]
        env = wo.build_stack(pz.pikazoo_v0.env(seed=seed, **wo.ENV_KW), W, stack)
        n_actions = env.action_space("player_1").n
        got = wo.run_session(env, n_actions, lambda f, a: po.synth_action(wo.ACTION_SEED, seed, f, a, n_actions))
        assert got["episodes"] == sess["episodes"], (name, seed)
        assert got["calls"] == sess["calls"] and got["sha256"] == sess["sha256"], (name, seed)
    # which of these ran fused
    raw = env.unwrapped
    fused = [f for _, f in raw._stack]
    assert all(fused) == (name == "canonical")
