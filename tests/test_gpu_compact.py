"""GPU: the byte-saving output formats — shared-row observations ([n][35]: player_1's row; player_2's is the same
values with the player blocks swapped, pikazoo_env.py:585-586) and the one-byte status (reward, terminated,
truncated) — through the device path, and the two-phase host-buffer call (pz_host_step_begin / _end) that ships
them over PCIe: 73 B per env-step instead of 305 B. Everything is compared with the oracle, bit for bit."""

import ctypes

import numpy as np
import pytest
import torch

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy

pytestmark = pytest.mark.gpu


def _expand(shared: np.ndarray) -> np.ndarray:
    """[n, 35] -> the reference's [n, 2, 35]"""
    from pikazoo_b200.vec_env import PLAYER2_INDEX

    return np.stack([shared, shared[:, list(PLAYER2_INDEX)]], axis=1)


def test_player2_index_is_the_reference_permutation(cuda_lib):
    from pikazoo_b200.vec_env import PLAYER2_INDEX

    assert [cuda_lib.pz_obs_player2_index(k) for k in range(35)] == list(PLAYER2_INDEX)
    assert cuda_lib.pz_obs_player2_index(35) == -1 and cuda_lib.pz_obs_player2_index(-1) == -1
    orc = po.OracleVecEnv(64, seed=3)
    obs = orc.reset()
    for t in range(50):
        obs, _, _ = orc.step(synth_actions_numpy(1, 0, 64, t, 18))
    assert np.array_equal(_expand(obs[:, 0]), obs)  # the oracle's player_2 rows ARE the permutation


@pytest.mark.parametrize("dtype", [torch.int16, torch.int32])
@pytest.mark.parametrize("n,cfg", [
    (4096 + 77, dict(winning_score=2, serve="random")),
    (1000, dict(winning_score=3, serve="alternate", is_player1_computer=True, is_player2_computer=True)),
    (31, dict(winning_score=1, serve="winner", is_player2_computer=True, max_episode_frames=60)),
])
def test_shared_rows_and_status_byte(cuda_lib, dtype, n, cfg):
    import pikazoo_b200

    env = pikazoo_b200.PikaVecEnv(n, seed=8, obs_dtype=dtype, obs_layout="shared", status=True, **cfg)
    orc = po.OracleVecEnv(n, seed=8, **cfg)
    assert env.obs.shape == (n, 35)
    assert np.array_equal(_expand(env.reset().cpu().numpy().astype(np.int32)), orc.reset())
    seen = 0
    for t in range(400):
        a = synth_actions_numpy(4, 0, n, t, 18)
        obs, rew, done = env.step(torch.from_numpy(a).cuda())
        o_obs, o_rew, o_done = orc.step(a)
        assert np.array_equal(_expand(obs.cpu().numpy().astype(np.int32)), o_obs), t
        st = env.status.cpu().numpy()
        assert np.array_equal((st & 3).astype(np.int64) - 1, o_rew[:, 0].astype(np.int64)), t
        assert np.array_equal(o_rew[:, 1], -o_rew[:, 0])
        assert np.array_equal((st >> 2) & 1, o_done), t
        assert np.array_equal((st >> 3) & 1, orc.truncated), t
        assert np.array_equal(rew.cpu().numpy(), o_rew.astype(np.float32))
        seen += int(o_done.sum())
    assert seen > 0 or cfg.get("is_player1_computer")  # (computer-vs-computer games outlast 400 frames)
    # reset / rollout / pz_observe emit the same layout
    env.rollout(16, actions="synth", action_seed=9, write_obs=True)
    orc.rollout(16, action_mode=1, action_seed=9, frame0=400)
    assert np.array_equal(env.export_state().cpu().numpy(), orc.state)
    assert np.array_equal(_expand(env.obs.cpu().numpy().astype(np.int32)), orc.current_obs())


def test_shared_layout_is_integer_only(cuda_lib):
    import pikazoo_b200

    with pytest.raises(TypeError):
        pikazoo_b200.PikaVecEnv(64, obs_layout="shared", obs_dtype=torch.float32)
    cfg = pikazoo_b200.make_config()
    cfg.obs_layout, cfg.obs_dtype = 2, 4
    assert cuda_lib.pz_reset(ctypes.c_void_p(16), 4, ctypes.byref(cfg), None, None) == -2


@pytest.mark.parametrize("n,chunks", [(5000, 4), (300_000, 0), (100, 1)])
def test_host_path_compact_two_phase(cuda_lib, n, chunks):
    """pz_host_step_begin / _end with int16 shared rows + status bytes + uint8 actions == the oracle; the call
    returns before the work is done (the buffers are only valid after _end), a second _begin in flight is refused."""
    import pikazoo_b200
    from pikazoo_b200 import _lib

    kw = dict(winning_score=3, serve="random", is_player2_computer=True)
    cfg = pikazoo_b200.make_config(obs_dtype=torch.int16, obs_layout="shared", action_dtype=torch.uint8, **kw)
    L = _lib.load()
    ctx = ctypes.c_void_p()
    _lib.check(L.pz_host_create(ctypes.byref(ctx), n, ctypes.byref(cfg), 77, 0, chunks), "pz_host_create")
    orc = po.OracleVecEnv(n, seed=77, **kw)
    obs_h = [torch.zeros((n, 35), dtype=torch.int16).pin_memory() for _ in range(2)]
    st_h = [torch.zeros((n,), dtype=torch.uint8).pin_memory() for _ in range(2)]
    _lib.check(L.pz_host_reset(ctx, obs_h[0].data_ptr()), "pz_host_reset")
    assert np.array_equal(_expand(obs_h[0].numpy().astype(np.int32)), orc.reset())
    steps = 120 if n > 100_000 else 300
    for t in range(steps):
        a = synth_actions_numpy(6, 0, n, t, 18)
        a_h = torch.from_numpy(a.astype(np.uint8)).pin_memory()
        b = t & 1
        _lib.check(L.pz_host_step_begin(ctx, a_h.data_ptr(), obs_h[b].data_ptr(), None, None, st_h[b].data_ptr()),
                   "pz_host_step_begin")
        assert L.pz_host_step_begin(ctx, a_h.data_ptr(), obs_h[b].data_ptr(), None, None, None) == -1  # in flight
        o_obs, o_rew, o_done = orc.step(a)  # the host is free while the GPU and the link work
        _lib.check(L.pz_host_step_end(ctx), "pz_host_step_end")
        assert np.array_equal(_expand(obs_h[b].numpy().astype(np.int32)), o_obs), t
        st = st_h[b].numpy()
        assert np.array_equal((st & 3).astype(np.int64) - 1, o_rew[:, 0].astype(np.int64))
        assert np.array_equal((st >> 2) & 1, o_done)
    assert L.pz_host_step_end(ctx) == -1  # nothing in flight
    stats = (ctypes.c_int64 * 16)()
    _lib.check(L.pz_host_stats(ctx, stats), "pz_host_stats")
    assert stats[0] == n * steps
    L.pz_host_destroy(ctx)


_SHAPED = ((0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4), 216, 176)


@pytest.mark.parametrize("n,chunks,threads,obs_dt,rew_dt,kw", [
    (5003, 4, 3, torch.int32, torch.float32, dict(winning_score=2, serve="random")),
    (300_000, 0, 0, torch.int32, torch.float32, dict(winning_score=1, serve="winner")),
    (1000, 2, 2, torch.int16, torch.float64, dict(winning_score=3, serve="alternate", is_player2_computer=True)),
    (777, 1, 5, torch.int32, torch.float64, dict(winning_score=2, serve="winner", reward_by_ball_position=_SHAPED)),
    (9, 1, 16, torch.int32, torch.float32, dict(winning_score=1, serve="winner", max_episode_frames=40)),
])
def test_host_path_wire_compact_is_transparent(cuda_lib, n, chunks, threads, obs_dt, rew_dt, kw):
    """pz_host_set_wire(PZ_WIRE_COMPACT): the caller's arrays keep the reference's dtypes and layout (int32 obs
    [n,2,35], reward [n,2], done [n]) while 71 B per env cross the link; host threads rebuild them. Every array ==
    the oracle's on every step (shaped rewards travel natively), switching the mode mid-run changes nothing."""
    import pikazoo_b200
    from pikazoo_b200 import _lib

    cfg = pikazoo_b200.make_config(obs_dtype=obs_dt, reward_dtype=rew_dt, **kw)
    L = _lib.load()
    ctx = ctypes.c_void_p()
    _lib.check(L.pz_host_create(ctypes.byref(ctx), n, ctypes.byref(cfg), 31, 0, chunks), "pz_host_create")
    _lib.check(L.pz_host_set_wire(ctx, 1, threads), "pz_host_set_wire")
    orc = po.OracleVecEnv(n, seed=31, **kw)
    np_obs = np.int32 if obs_dt == torch.int32 else np.int16
    obs_h = torch.full((n, 2, 35), -7, dtype=obs_dt).pin_memory()
    rew_h = torch.full((n, 2), 9.0, dtype=rew_dt).pin_memory()
    done_h = torch.full((n,), 9, dtype=torch.uint8).pin_memory()
    st_h = torch.full((n,), 255, dtype=torch.uint8).pin_memory()
    _lib.check(L.pz_host_reset(ctx, obs_h.data_ptr()), "pz_host_reset")
    assert np.array_equal(obs_h.numpy(), orc.reset().astype(np_obs))
    steps = 100 if n > 100_000 else 300
    dones = 0
    for t in range(steps):
        if t == steps // 2:  # back to the native format for a while: the state lives in the context, not the format
            _lib.check(L.pz_host_set_wire(ctx, 0, 0), "pz_host_set_wire")
        if t == steps // 2 + 20:
            _lib.check(L.pz_host_set_wire(ctx, 1, threads), "pz_host_set_wire")
        a = synth_actions_numpy(6, 0, n, t, 18).astype(np.int32)
        a_h = torch.from_numpy(a).pin_memory()
        _lib.check(L.pz_host_step_begin(ctx, a_h.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr(),
                                        st_h.data_ptr() if t % 3 == 0 else None), "pz_host_step_begin")
        o_obs, o_rew, o_done = orc.step(a)
        _lib.check(L.pz_host_step_end(ctx), "pz_host_step_end")
        assert np.array_equal(obs_h.numpy(), o_obs.astype(np_obs)), t
        assert np.array_equal(rew_h.numpy().view(np.uint8), o_rew.astype(rew_h.numpy().dtype).view(np.uint8)), t
        assert np.array_equal(done_h.numpy(), o_done.astype(np.uint8)), t
        if t % 3 == 0:
            assert np.array_equal((st_h.numpy() >> 2) & 1, o_done), t
        dones += int(o_done.sum())
    assert dones > 0
    # pz_host_step (one call) goes through the same path
    a = synth_actions_numpy(6, 0, n, steps, 18).astype(np.int32)
    a_h = torch.from_numpy(a).pin_memory()
    _lib.check(L.pz_host_step(ctx, a_h.data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr()), "pz_host_step")
    o_obs, o_rew, o_done = orc.step(a)
    assert np.array_equal(obs_h.numpy(), o_obs.astype(np_obs)) and np.array_equal(done_h.numpy(), o_done.astype(np.uint8))
    L.pz_host_destroy(ctx)


def test_host_wire_needs_integer_env_major_rows(cuda_lib):
    import pikazoo_b200
    from pikazoo_b200 import _lib

    L = _lib.load()
    for kw in (dict(obs_dtype=torch.float32, normalize_observation=True), dict(obs_dtype=torch.int16, obs_layout="shared")):
        cfg = pikazoo_b200.make_config(**kw)
        ctx = ctypes.c_void_p()
        _lib.check(L.pz_host_create(ctypes.byref(ctx), 256, ctypes.byref(cfg), 1, 0, 1), "pz_host_create")
        assert L.pz_host_set_wire(ctx, 1, 2) == -2
        assert L.pz_host_set_wire(ctx, 7, 2) == -1
        L.pz_host_destroy(ctx)
