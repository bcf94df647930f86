"""GPU: out-of-bounds writes, found with our own guard bands (compute-sanitizer is not available on the GPU
pool). Every output buffer of every kernel family is a slice of a larger allocation whose guard bands hold a
pattern; after the launches the bands must be untouched and every element inside must have been written. Sizes
cover single envs, ragged warps, ragged CTAs / policy tiles and the full-CTA fast paths (bulk copies, staged
feature-major rows)."""

import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on either side
PATTERN = 0xA5


class Guarded:
    """`nbytes` of device memory between two guard bands; 256-byte aligned payload."""

    def __init__(self, nbytes: int, fill: int = 0x5A):
        self.nbytes = int(nbytes)
        self.pad = (-self.nbytes) % 256
        self.raw = torch.full((GUARD + self.nbytes + self.pad + GUARD,), PATTERN, dtype=torch.uint8, device="cuda")
        self.payload = self.raw[GUARD:GUARD + self.nbytes]
        self.payload.fill_(fill)

    @property
    def ptr(self) -> int:
        return self.payload.data_ptr()

    def view(self, dtype, shape):
        return self.payload.view(dtype).view(shape)

    def check(self, what):
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.nbytes:]
        assert bool((lo == PATTERN).all()) and bool((hi == PATTERN).all()), f"{what}: guard band overwritten"


_DT = {torch.int32: 4, torch.int16: 2, torch.float32: 4, torch.float16: 2, torch.bfloat16: 2, torch.float64: 8}


@pytest.mark.parametrize("n", [1, 33, 128, 1000, 4096, 4096 + 24])
@pytest.mark.parametrize("layout", ["env_major", "feature_major"])
def test_step_reset_rollout_stay_inside_their_buffers(cuda_lib, n, layout):
    import pikazoo_b200
    from pikazoo_b200 import _lib

    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(n)
    for obs_dtype, normalize in ((torch.int32, False), (torch.int16, False), (torch.bfloat16, True),
                                 (torch.float32, True), (torch.float64, True)):
        for ai in (False, True):
            rows = 40 if layout == "feature_major" else 35
            cfg = pikazoo_b200.make_config(winning_score=2, serve="random", is_player1_computer=ai,
                                           is_player2_computer=ai, obs_dtype=obs_dtype, normalize_observation=normalize,
                                           obs_layout=layout, obs_feature_rows=rows, landing_tables=False,
                                           max_episode_frames=40, reward_dtype=torch.float64)
            esz = _DT[obs_dtype]
            state = Guarded(_lib.STATE_WORDS * n * 4, fill=0)
            obs = Guarded(2 * rows * n * esz)
            reward = Guarded(n * 2 * 8)
            done = Guarded(n)
            trunc = Guarded(n)
            ep_ret = Guarded(n * 2 * 8, fill=0)
            ep_len = Guarded(n * 4)
            stats = Guarded(_lib.NUM_STATS * 8, fill=0)
            ep = _lib.PzEpisodeIo(ep_ret.ptr, ep_len.ptr, trunc.ptr)
            _lib.check(L.pz_seed(state.ptr, n, 7, 0, None))
            _lib.check(L.pz_reset_ex(state.ptr, n, ctypes.byref(cfg), obs.ptr, ctypes.byref(ep), None))
            for t in range(5):
                a = torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32)
                _lib.check(L.pz_step_ex(state.ptr, n, ctypes.byref(cfg), a.data_ptr(), obs.ptr, reward.ptr, done.ptr,
                                        stats.ptr, ctypes.byref(ep), None))
            _lib.check(L.pz_rollout(state.ptr, n, ctypes.byref(cfg), 6, _lib.ACTIONS_SYNTH, 1, 0, 5, obs.ptr, stats.ptr,
                                    None))
            torch.cuda.synchronize()
            label = (n, layout, str(obs_dtype), ai)
            for name, buf in (("state", state), ("obs", obs), ("reward", reward), ("done", done), ("truncated", trunc),
                              ("episode_return", ep_ret), ("episode_length", ep_len), ("stats", stats)):
                buf.check((name, label))
            # everything that should be written was written (the fill pattern 0x5A.. is not a valid value anywhere)
            assert bool((done.payload <= 1).all()) and bool((trunc.payload <= 1).all()), label
            assert bool((reward.view(torch.float64, (n, 2)).abs() <= 1).all()), label
            o = obs.view(obs_dtype, (2, rows, n) if layout == "feature_major" else (n, 2, 35))
            body = o[:, :35, :] if layout == "feature_major" else o
            assert bool((body.double().abs() <= 433).all()), label  # 0x5A5A.. would be ~2e4 (i16) / 1e9 (i32) / 1.5e16 (f32)
            if layout == "feature_major":  # rows 35.. belong to the caller
                assert bool((o[:, 35:, :].contiguous().view(torch.uint8) == 0x5A).all()), label


@pytest.mark.parametrize("n", [1, 16, 127, 128, 1000, 4096 + 8, 100_003])
def test_policy_kernel_stays_inside_its_buffers(cuda_lib, n):
    from pikazoo_b200 import _lib
    from pikazoo_b200.policy import MLPPolicy

    L = _lib.load()
    policy = MLPPolicy(device="cuda", seed=2)
    obs = Guarded(2 * 40 * n * 2, fill=0)
    obs.view(torch.bfloat16, (2, 40, n)).copy_(torch.rand((2, 40, n), device="cuda"))
    for code, width in ((_lib.ACT_U8, 1), (_lib.ACT_I32, 4), (_lib.ACT_I64, 8)):
        actions = Guarded(n * 2 * width)
        logits = Guarded(n * 2 * 18 * 4)
        _lib.check(L.pz_policy_mlp_act(obs.ptr, n, n, 40, policy.w1.data_ptr(), 72, 40, policy.w2.data_ptr(), 18, 72,
                                       3, 1, 0, actions.ptr, code, 0, logits.ptr, None))
        torch.cuda.synchronize()
        for name, buf in (("obs", obs), ("actions", actions), ("logits", logits)):
            buf.check((name, n, code))
        dt = {1: torch.uint8, 4: torch.int32, 8: torch.int64}[width]
        a = actions.view(dt, (n, 2))
        assert bool((a >= 0).all()) and bool((a < 18).all())
        assert bool(torch.isfinite(logits.view(torch.float32, (n, 2, 18))).all())
