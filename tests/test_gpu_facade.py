"""GPU: the reference-shaped single-env API (pikazoo_v0.env + wrappers). Mirrors the
reference's own tests (tests/env/test_env.py observation symmetry; the dict protocol that
tests/test_parallel_api.py exercises through pettingzoo) and checks values against the oracle."""

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle.synth import synth_action

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pikazoo_v0(cuda_lib):
    from pikazoo_b200 import pikazoo_v0 as m

    return m


def _divide_and_assert(observations):
    # reference tests/env/test_env.py:16-21
    p1a, p2a = observations["player_1"][0:13], observations["player_1"][13:26]
    p2b, p1b = observations["player_2"][0:13], observations["player_2"][13:26]
    assert np.all(p1a == p1b) and np.all(p2a == p2b)


def test_env_observation_symmetry(pikazoo_v0):
    # reference tests/env/test_env.py:7-14 (AI vs AI, NOOP actions) with a seed that terminates
    env = pikazoo_v0.env(winning_score=3, is_player1_computer=True, is_player2_computer=True, render_mode=None, seed=0)
    observations, infos = env.reset()
    _divide_and_assert(observations)
    frames = 0
    while env.agents and frames < 20000:
        actions = {agent: 0 for agent in env.agents}
        observations, rewards, terminations, truncations, infos = env.step(actions)
        _divide_and_assert(observations)
        frames += 1
    assert not env.agents and max(env.scores) == 3
    with pytest.raises(IndexError):
        env.step({"player_1": 0, "player_2": 0})


def test_parallel_api_protocol_and_values(pikazoo_v0):
    seed = 42
    env = pikazoo_v0.parallel_env(winning_score=2, serve="random", seed=seed)
    orc = po.OracleVecEnv(1, seed=seed, autoreset=False, winning_score=2, serve="random")
    obs, infos = env.reset()
    assert set(obs) == {"player_1", "player_2"} and obs["player_1"].shape == (35,)
    assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), orc.reset()[0])
    assert infos["player_1"] == {"score": [0, 0]}
    t = 0
    while env.agents:
        a = {ag: synth_action(1, 0, t, i) for i, ag in enumerate(env.agents)}
        for ag in env.agents:
            assert env.action_space(ag).contains(a[ag])
        obs, rew, term, trunc, infos = env.step(a)
        o, r, d = orc.step(np.array([[a["player_1"], a["player_2"]]], dtype=np.int32))
        assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), o[0])
        assert env.observation_space("player_1").contains(obs["player_1"].astype(np.int32))
        assert [rew["player_1"], rew["player_2"]] == [int(r[0, 0]), int(r[0, 1])]
        assert isinstance(rew["player_1"], int) and rew["player_1"] == -rew["player_2"]
        assert term == {"player_1": bool(d[0]), "player_2": bool(d[0])}
        assert trunc == {"player_1": False, "player_2": False}
        assert infos["player_1"]["score"] == list(orc.state[0, 37:39])
        assert np.array_equal(env.state_words(), orc.state[0, :52])
        t += 1
    assert max(env.scores) == 2
    # reset() on the same object carries state over exactly like the reference's
    obs, _ = env.reset()
    orc.cfg.autoreset = 1
    assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), orc.reset()[0])
    with pytest.raises(IndexError):
        env.step({"player_1": 18, "player_2": 0})


def test_wrappers_fused_in_kernel(pikazoo_v0):
    from pikazoo_b200.wrappers import RewardByBallPosition, SimplifyAction

    add = (0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4)
    env = RewardByBallPosition(SimplifyAction(pikazoo_v0.env(winning_score=2, seed=7)), add, 200, 150)
    orc = po.OracleVecEnv(1, seed=7, autoreset=False, winning_score=2, simplify_action=True,
                          reward_by_ball_position=(add, 200, 150))
    obs, _ = env.reset()
    orc.reset()
    t = 0
    while env.agents:
        a = {ag: synth_action(2, 0, t, i, 13) for i, ag in enumerate(env.agents)}
        obs, rew, term, _, _ = env.step(a)
        o, r, d = orc.step(np.array([[a["player_1"], a["player_2"]]], dtype=np.int32))
        assert np.array_equal(np.stack([obs["player_1"], obs["player_2"]]), o[0])
        assert rew["player_1"] == r[0, 0] and rew["player_2"] == r[0, 1]  # exact doubles
        t += 1
    with pytest.raises(IndexError):
        env.reset()
        env.step({"player_1": 13, "player_2": 0})


@pytest.mark.parametrize("n", [1, 100, 4096])
def test_host_mapped_buffers_equal_device_buffers(cuda_lib, n):
    """PikaVecEnv(host_mapped=True): every per-call buffer is pinned host memory addressed by the kernels
    directly (the facade's mode; full warps leave through the bulk copy, ragged tails through vector stores).
    Same trajectories, bit for bit, as the device-resident env, with computer players and statistics on."""
    import torch

    import pikazoo_b200

    kw = dict(seed=77, winning_score=3, serve="random", is_player2_computer=True, record_episode_statistics=True,
              max_episode_frames=400)
    dev = pikazoo_b200.PikaVecEnv(n, **kw)
    host = pikazoo_b200.PikaVecEnv(n, host_mapped=True, **kw)
    assert host.obs.device.type == "cpu" and host.obs.is_pinned()
    o_d, o_h = dev.reset(), host.reset()
    torch.cuda.synchronize()
    assert torch.equal(o_d.cpu(), o_h)
    g = torch.Generator().manual_seed(5)
    a_h = torch.zeros((n, 2), dtype=torch.int32).pin_memory()
    for t in range(600):
        a_h.copy_(torch.randint(0, 18, (n, 2), generator=g, dtype=torch.int32))
        od, rd, dd = dev.step(a_h.cuda())
        oh, rh, dh = host.step(a_h)
        torch.cuda.synchronize()
        assert torch.equal(od.cpu(), oh) and torch.equal(rd.cpu(), rh) and torch.equal(dd.cpu(), dh), t
        assert torch.equal(dev.truncated.cpu(), host.truncated)
    assert torch.equal(dev.state.cpu(), host.state.cpu())  # (the packed state stays in device memory in both modes)
    assert torch.equal(dev.episode_return.cpu(), host.episode_return)
    assert torch.equal(dev.export_state().cpu(), host.export_state().cpu())
    assert dev.stats_dict() == host.stats_dict()
    with pytest.raises(ValueError):
        host.step(torch.zeros((n, 2), dtype=torch.int32))  # not pinned
