"""CPU: host-side logic — sharding, config mirroring, facade argument behaviour, and the
world_size-2 statistics all-reduce over gloo."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pikazoo_b200
from pikazoo_b200 import make_config, shard_range
from pikazoo_b200 import pikazoo_v0
from pikazoo_b200.wrappers import RewardByBallPosition, SimplifyAction


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1000, 1 << 20, 16 * (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_make_config_mirrors_reference_arguments():
    c = make_config(winning_score=5, serve="random", is_player2_computer=True, simplify_action=True,
                    reward_by_ball_position=((0.1,) * 8, 200, 100), action_dtype=torch.int64,
                    reward_dtype=torch.float64)
    assert (c.winning_score, c.serve, c.is_player1_computer, c.is_player2_computer) == (5, 2, 0, 1)
    assert (c.simplify_action, c.reward_by_ball_position, c.x_line, c.y_line) == (1, 1, 200, 100)
    assert (c.action_dtype, c.reward_dtype) == (1, 1)
    with pytest.raises(AssertionError):  # reference: assert serve in (...), pikazoo_env.py:104
        make_config(serve="loser")
    with pytest.raises(ValueError):
        make_config(winning_score=0)
    with pytest.raises(AssertionError):  # reference: assert len(additional_reward) == 8
        make_config(reward_by_ball_position=((1, 2, 3), 216, 176))


def test_facade_surface_matches_reference():
    env = pikazoo_v0.env(winning_score=7, serve="alternate")
    assert pikazoo_v0.parallel_env is not None and pikazoo_v0.raw_env is type(env)
    assert env.possible_agents == ["player_1", "player_2"] and env.agents == env.possible_agents
    assert env.action_space("player_1").n == 18 and env.action_space("player_2").n == 18
    sp = env.observation_space("player_1")
    assert sp.shape == (35,) and sp.dtype == np.int32
    assert sp.low[0] == 32 and sp.high[0] == 400 and sp.low[33] == -124 and sp.high[26] == 432
    assert env.metadata["name"] == "pikazoo_v0"
    with pytest.raises(AssertionError):
        pikazoo_v0.env(serve="nobody")
    with pytest.raises(NotImplementedError):
        pikazoo_v0.env(render_mode="human")
    w = RewardByBallPosition(SimplifyAction(env), (1, 2, 3, 4, 5, 6, 7, 8))
    assert w.action_space("player_2").n == 13
    assert w.unwrapped is env and env._simplify_action and env._reward_by_ball_position[1:] == (216, 176)
    with pytest.raises(AssertionError):
        RewardByBallPosition(env, (1, 2, 3))


def test_no_cpu_fallback():
    with pytest.raises(pikazoo_b200.PikaLibraryError):
        pikazoo_b200.PikaVecEnv(4, device="cpu")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stats_worker(rank, world, port, total_envs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(total_envs, world, rank)
    # each rank contributes the statistics of its own shard: here a function of the global env ids
    ids = torch.arange(first, first + count, dtype=torch.int64)
    stats = torch.zeros(16, dtype=torch.int64)
    stats[0] = count
    stats[1] = (ids % 3 == 0).sum()
    stats[2] = ids.sum()
    pikazoo_b200.allreduce_stats(stats)
    out[rank] = stats.tolist()
    dist.destroy_process_group()


def test_stats_allreduce_world_size_2_gloo():
    total = 1001
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_stats_worker, args=(2, port, total, out), nprocs=2, join=True)
        ids = torch.arange(total)
        want = [total, int((ids % 3 == 0).sum()), int(ids.sum())]
        assert out[0][:3] == want and out[1][:3] == want


def test_allreduce_is_noop_without_process_group():
    s = torch.arange(16, dtype=torch.int64)
    assert pikazoo_b200.allreduce_stats(s) is None
    assert s.tolist() == list(range(16))


def test_policy_noise_and_sampling_restatements():
    """The numpy restatements the GPU tests check the fused policy kernel against: the counter-based noise is a
    pure function of its counters, differs between envs / agents / actions / steps, and is Gumbel-distributed (up
    to the constant ln ln 2); the packed-key arg-max is the plain arg-max except between keys equal in their upper
    27 bits, where positive keys prefer the lower action."""
    from pikazoo_b200.policy import gumbel_noise_reference, sample_reference

    a = gumbel_noise_reference(7, 3, 100, 4096, 18)
    assert a.shape == (4096, 2, 18) and a.dtype == np.float32 and np.isfinite(a).all()
    assert np.array_equal(a, gumbel_noise_reference(7, 3, 100, 4096, 18))
    assert np.array_equal(a[50:60], gumbel_noise_reference(7, 3, 150, 10, 18))  # keyed by the GLOBAL env index
    assert not np.array_equal(a, gumbel_noise_reference(7, 4, 100, 4096, 18))
    assert not np.array_equal(a, gumbel_noise_reference(8, 3, 100, 4096, 18))
    assert len(np.unique(a)) > 0.99 * a.size
    g = a.astype(np.float64) - np.log(np.log(2.0))  # standard Gumbel: mean 0.5772, variance pi^2 / 6
    assert abs(g.mean() - 0.5772) < 0.02 and abs(g.var() - np.pi ** 2 / 6) < 0.05
    # uniform logits: every action equally likely
    counts = np.bincount(sample_reference(np.zeros((4096, 2, 18), np.float32), a).ravel(), minlength=18)
    assert counts.min() > 0.8 * counts.mean() and counts.max() < 1.2 * counts.mean()
    logits = np.random.default_rng(0).normal(size=(1000, 2, 18)).astype(np.float32)
    got, plain = sample_reference(logits, None), logits.argmax(axis=-1)
    for e, ag in np.argwhere(got != plain):  # only between keys closer than the 5 truncated mantissa bits
        assert abs(logits[e, ag, got[e, ag]] - logits[e, ag, plain[e, ag]]) < 32 * np.spacing(np.float32(abs(logits[e, ag]).max()))
    assert (got != plain).sum() <= 3
    tie = np.zeros((1, 2, 18), np.float32)
    tie[0, :, 5] = tie[0, :, 11] = 2.0
    assert sample_reference(tie, None).tolist() == [[5, 5]]
    # the tcgen05 kernel's sampler: inversion of the cumulative distribution with the stream's first uniform
    from pikazoo_b200.policy import _counter_uniform, inverse_cdf_reference

    u = _counter_uniform(7, 3, 100, 4096, 1)
    assert u.shape == (4096, 2, 1) and u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    assert np.array_equal(u[..., 0], _counter_uniform(7, 3, 100, 4096, 18)[..., 0])  # slot 0 of the Gumbel stream
    lg = np.random.default_rng(1).normal(size=(100_000, 2, 18)).astype(np.float32)
    s = inverse_cdf_reference(lg, 7, 3, 100)
    assert s.shape == (100_000, 2) and s.min() == 0 and s.max() == 17
    assert np.array_equal(s[50:60], inverse_cdf_reference(lg[50:60], 7, 3, 150))  # keyed by the GLOBAL env index
    p = np.exp(lg.astype(np.float64)); p /= p.sum(axis=-1, keepdims=True)
    freq, want = np.bincount(s.ravel(), minlength=18) / s.size, p.mean(axis=(0, 1))
    assert np.abs(freq - want).max() < 5 * np.sqrt(want.max() / s.size)
    one_hot = np.full((4, 2, 18), -np.inf, np.float32)
    one_hot[:, :, 13] = 0.0
    assert (inverse_cdf_reference(one_hot, 1, 2, 3) == 13).all()  # zero-width intervals are never chosen
