"""GPU: the documentation is executable. The ctypes stub INTEGRATION.md §3 shows a maintainer of the
reference (host buffers, numpy only) is extracted from the document and run against the built library
and the oracle; and a restatement of what pettingzoo's `parallel_api_test` checks (the reference's own
tests/test_parallel_api.py) runs over the facade with and without wrappers."""

import os
import re

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_integration_md_stub_runs(cuda_lib):
    from pikazoo_b200 import _lib
    from tests.test_abi import integration_stub_source

    block = integration_stub_source()

    block = block.replace('ctypes.CDLL("libpikazoo_b200.so")', f'ctypes.CDLL("{_lib.LIB_PATH}")')
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)  # noqa: S102 - our own document
    n = 3000
    env = ns["VecPikaEnv"](n, seed=5, winning_score=3, serve="random", is_player2_computer=True)
    orc = po.OracleVecEnv(n, seed=5, winning_score=3, serve="random", is_player2_computer=True)
    assert np.array_equal(env.reset(), orc.reset())
    rng = np.random.default_rng(0)
    for t in range(200):
        a = rng.integers(0, 18, size=(n, 2), dtype=np.int32)
        obs, rew, done = env.step(a)
        o_obs, o_rew, o_done = orc.step(a)
        assert np.array_equal(obs, o_obs) and np.array_equal(rew, o_rew.astype(np.float32))
        assert np.array_equal(done, o_done.astype(bool))


def _parallel_api_cycles(env, cycles, n_actions):
    """the invariants pettingzoo.test.parallel_api_test asserts, restated (pettingzoo is not installed)"""
    assert hasattr(env, "possible_agents") and len(set(env.possible_agents)) == len(env.possible_agents) == 2
    obs, infos = env.reset()
    assert isinstance(obs, dict) and isinstance(infos, dict) and set(obs) == set(env.agents) == set(infos)
    episodes = 0
    for _ in range(cycles):
        live = list(env.agents)
        actions = {a: env.action_space(a).sample() for a in live}
        assert all(0 <= v < n_actions for v in actions.values())
        obs, rew, term, trunc, infos = env.step(actions)
        for d in (obs, rew, term, trunc, infos):
            assert isinstance(d, dict) and set(d) == set(live)
        for a in live:
            space = env.observation_space(a)
            assert obs[a].shape == space.shape
            assert isinstance(term[a], bool) and isinstance(trunc[a], bool) and isinstance(infos[a], dict)
            assert np.isfinite(float(rew[a]))
        assert term[live[0]] == term[live[1]]
        if term[live[0]]:
            assert env.agents == []  # terminated agents leave the env (pikazoo_env.py:237-238)
            episodes += 1
            obs, infos = env.reset()
            assert set(env.agents) == set(env.possible_agents)
        else:
            assert env.agents == live
    return episodes


def test_parallel_api_conformance(cuda_lib):
    from pikazoo_b200 import pikazoo_v0
    from pikazoo_b200 import wrappers as W

    assert _parallel_api_cycles(pikazoo_v0.env(winning_score=1, seed=1), 1500, 18) >= 1
    env = pikazoo_v0.env(winning_score=1, serve="random", seed=2)
    env = W.RecordEpisodeStatistics(W.NormalizeObservation(W.RewardInNormalState(
        W.RewardByBallPosition(W.SimplifyAction(env), (0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4)), 0.01)))
    assert _parallel_api_cycles(env, 1500, 13) >= 1
    assert pikazoo_v0.env().metadata["name"] == "pikazoo_v0"
    with pytest.raises(AssertionError):
        pikazoo_v0.env(serve="loser")  # pikazoo_env.py:104
