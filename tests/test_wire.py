"""pz_wire_expand — the host-side half of the compact wire format of the host-buffer path (pure host code, so it is
checked here without a GPU): player_1's int16 row + status byte -> the reference's obs [n][2][35] (player_2's row is
[own block | opponent's block | ball], pikazoo/env/pikazoo_env.py:585-586), reward [n][2], done [n]."""

import ctypes

import numpy as np
import pytest

from pikazoo_b200 import _lib


def _expand_numpy(rows, status, obs_dtype, rew_dtype):
    p1 = rows.astype(obs_dtype)
    p2 = np.concatenate([p1[:, 13:26], p1[:, 0:13], p1[:, 26:35]], axis=1)
    base = (status & 3).astype(np.int64) - 1
    rew = np.stack([base, -base], axis=1).astype(rew_dtype)
    return np.stack([p1, p2], axis=1), rew, ((status >> 2) & 1).astype(np.uint8)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 64, 1001, 40_003])
@pytest.mark.parametrize("obs_dtype,code", [(np.int32, 0), (np.int16, 1)])
@pytest.mark.parametrize("misalign", [0, 1])
def test_wire_expand_matches_numpy(n, obs_dtype, code, misalign):
    L = _lib.load()
    rng = np.random.default_rng(n * 7 + code)
    rows = rng.integers(-32768, 32768, size=(n, 35), dtype=np.int16)
    status = rng.integers(0, 16, size=(n,), dtype=np.uint8)
    status = (status & 0xC) | rng.integers(0, 3, size=(n,), dtype=np.uint8)  # reward code 0..2
    # outputs inside guard bands, optionally off the 32-byte alignment the streaming stores want
    pad = 64
    raw = np.full(n * 70 + 2 * pad, 0x5A5A, dtype=obs_dtype)
    obs = raw[pad + misalign: pad + misalign + n * 70].reshape(n, 2, 35)
    for rew_dtype, rcode in ((np.float32, 0), (np.float64, 1)):
        raw[:] = 0x5A5A
        rew = np.full((n, 2), 9.0, dtype=rew_dtype)
        done = np.full((n,), 9, dtype=np.uint8)
        rc = L.pz_wire_expand(_ptr(rows), _ptr(status), n, code, _ptr(obs), rcode, _ptr(rew), _ptr(done))
        assert rc == 0
        e_obs, e_rew, e_done = _expand_numpy(rows, status, obs_dtype, rew_dtype)
        assert np.array_equal(obs, e_obs)
        assert np.array_equal(rew, e_rew)
        assert np.array_equal(done, e_done)
        assert (raw[:pad + misalign] == 0x5A5A).all() and (raw[pad + misalign + n * 70:] == 0x5A5A).all()


def test_wire_expand_arguments():
    L = _lib.load()
    rows = np.zeros((4, 35), np.int16)
    status = np.zeros((4,), np.uint8)
    obs = np.zeros((4, 2, 35), np.float32)
    assert L.pz_wire_expand(_ptr(rows), _ptr(status), 4, 2, _ptr(obs), 0, None, None) == -2  # float rows
    assert L.pz_wire_expand(None, _ptr(status), 4, 0, _ptr(obs), 0, None, None) == -1
    assert L.pz_wire_expand(_ptr(rows), None, 4, 0, None, 0, _ptr(obs), None) == -1  # reward needs status
    assert L.pz_wire_expand(_ptr(rows), None, 4, 0, None, 0, None, None) == 0  # nothing asked for
