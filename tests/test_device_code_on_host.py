"""The DEVICE headers (pz_state.cuh / pz_rng.cuh / pz_physics.cuh) compiled for the host with one
lane per warp (tests/emul/pz_emul.cpp) against the oracle, without a GPU: packing round trips,
PCG64 seeding and draws, the fast-forwarded trajectory simulations over a domain much wider than
play reaches, and lock-step games for every AI mask. The kernels proper (launch geometry, bulk-copy
output, warp-collective scheduling, statistics) are covered by the -m gpu tests."""

import ctypes

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle.synth import synth_actions_numpy
from tests.emul.build import build

LIB = build()
pytestmark = pytest.mark.skipif(LIB is None, reason="CUDA headers not found")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def emul():
    L = ctypes.CDLL(LIB)
    L.emul_simulate.restype = ctypes.c_int
    L.emul_synth_action.restype = ctypes.c_int
    L.emul_synth_action.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32]
    L.emul_seed.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64]
    L.emul_import.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    L.emul_export.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    L.emul_simulate_many.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.emul_step.argtypes = [ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 4 + [ctypes.c_int]
    return L


class EmulVecEnv:
    def __init__(self, L, n, seed, winning_score=15, serve="winner", is_player1_computer=False,
                 is_player2_computer=False, simplify_action=False, autoreset=True):
        self.L, self.n = L, n
        self.packed = np.zeros(17 * n, dtype=np.int32)
        self.args = (winning_score, po.SERVE_CODES[serve], int(is_player1_computer) | 2 * int(is_player2_computer),
                     int(simplify_action), int(autoreset))
        self.obs = np.zeros((n, 2, 35), dtype=np.int32)
        self.base = np.zeros(n, dtype=np.int32)
        self.done = np.zeros(n, dtype=np.uint8)
        L.emul_seed(_p(self.packed), n, seed)

    def _call(self, actions, reset):
        a = None if actions is None else np.ascontiguousarray(actions, dtype=np.int32)
        self.L.emul_step(_p(self.packed), self.n, *self.args, None if a is None else _p(a), _p(self.obs),
                         _p(self.base), _p(self.done), int(reset))

    def reset(self):
        self._call(None, True)
        return self.obs

    def step(self, actions):
        self._call(actions, False)
        return self.obs, self.base, self.done

    def export(self):
        out = np.zeros((self.n, 53), dtype=np.int32)
        self.L.emul_export(_p(self.packed), self.n, _p(out))
        return out

    def load(self, unpacked):
        self.L.emul_import(_p(self.packed), self.n, _p(np.ascontiguousarray(unpacked, dtype=np.int32)))


def test_seeding_matches_oracle(emul):
    n = 257
    for base in (0, 1, 12345, 2**32 - 3, 2**63 + 11, 2**64 - n):
        e = EmulVecEnv(emul, n, base)
        o = po.OracleVecEnv(n, seed=base)
        assert np.array_equal(e.export(), o.state)


@pytest.mark.parametrize("power", [False, True])
def test_fast_forwarded_simulation_equals_plain_loop(emul, power):
    """simulate_landing_x (closed-form skips of free flight) == the reference loops, on the whole
    reachable domain and far beyond it (negative y after net bounces, |yv| up to 700, loop limit)."""
    rng = np.random.default_rng(7 + power)
    n = 400_000
    q = np.empty((n, 4), dtype=np.int32)
    q[:, 0] = rng.integers(0, 433, n)
    q[:, 1] = rng.integers(-512, 253, n)
    q[:, 2] = rng.integers(-20, 21, n)
    q[:, 3] = rng.integers(-700, 701, n)
    # a third of the cases inside the domain play actually reaches, dense around the net and walls
    m = n // 3
    q[:m, 1] = rng.integers(0, 253, m)
    q[:m, 3] = rng.integers(-60, 61, m)
    q[m:2 * m, 0] = rng.integers(180, 253, m)
    q[m:2 * m, 1] = rng.integers(150, 253, m)
    if power:
        q[:, 2] = rng.choice(np.array([-20, -10, 10, 20], dtype=np.int32), n)
    out = np.zeros(n, dtype=np.int32)
    emul.emul_simulate_many(n, _p(q), int(power), _p(out))
    ref = po.simulate_many(q, power)
    bad = np.nonzero(out != ref)[0]
    assert len(bad) == 0, (q[bad[:5]].tolist(), out[bad[:5]].tolist(), ref[bad[:5]].tolist())


def test_simulation_exhaustive_slice(emul):
    """every (x, y) for a band of velocities, both loops"""
    xs, ys = np.meshgrid(np.arange(0, 433, dtype=np.int32), np.arange(0, 253, dtype=np.int32), indexing="ij")
    for xv, yv in [(0, 0), (20, -31), (-20, 40), (7, 1), (-13, -90), (10, 262), (-20, -262), (3, 101)]:
        q = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, xv, np.int32), np.full(xs.size, yv, np.int32)], 1)
        q = np.ascontiguousarray(q, dtype=np.int32)
        for power in (0, 1):
            out = np.zeros(len(q), dtype=np.int32)
            emul.emul_simulate_many(len(q), _p(q), power, _p(out))
            assert np.array_equal(out, po.simulate_many(q, bool(power))), (xv, yv, power)


CONFIGS = {
    "random18": dict(),
    "ai_vs_ai": dict(is_player1_computer=True, is_player2_computer=True),
    "ai_p1_alternate": dict(is_player1_computer=True, serve="alternate", winning_score=7),
    "ai_p2_random_serve": dict(is_player2_computer=True, serve="random", winning_score=5),
    "simplify_ws3": dict(simplify_action=True, winning_score=3, serve="random"),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_lockstep_with_oracle(emul, name):
    cfg = CONFIGS[name]
    n, steps, seed = 192, 2500, 30_000  # seed 30000 / env 3214 was the fast-forward ceiling bug's witness
    e = EmulVecEnv(emul, n, seed, **cfg)
    o = po.OracleVecEnv(n, seed=seed, **cfg)
    assert np.array_equal(e.reset(), o.reset())
    n_actions = 13 if cfg.get("simplify_action") else 18
    for t in range(steps):
        a = synth_actions_numpy(seed + 77, 0, n, t, n_actions)
        obs, base, done = e.step(a)
        o_obs, o_rew, o_done = o.step(a)
        assert np.array_equal(obs, o_obs), f"obs differ at step {t}"
        assert np.array_equal(base, o_rew[:, 0].astype(np.int32)) and np.array_equal(done, o_done), t
        if t % 50 == 0 or t == steps - 1:
            assert np.array_equal(e.export(), o.state), f"state differs at step {t}"


def test_known_witness_of_the_ceiling_bug(emul):
    """State (from a 4,096-env GPU run) whose power-hit candidate flies above the ceiling (y = -10):
    the computer must find the (x_direction 1, y_direction 1) hit and move right."""
    prev = [178, 111, 3, 2, 2, 0, 1, 1, -2, 0, 2, 0, 0, 248, 228, 16, 2, 0, 2, 1, 1, -2, 0, 1, 0, 0, 213, 243, 10,
            131, 203, 113, 223, 179, 1, 213, 203, 0, 0, 0, 0, 0, 2047048881, -470378562, -633581333, 1788586230,
            -1227392537, 1011126410, 2026769623, -172115264, 0, 264967842, 249]
    cfg = dict(winning_score=5, serve="random", is_player2_computer=True)
    e = EmulVecEnv(emul, 1, 0, **cfg)
    e.load(np.array([prev], dtype=np.int32))
    o = po.OracleVecEnv(1, seed=0, **cfg)
    o.state[:] = np.array([prev], dtype=np.int32)
    a = np.array([[15, 3]], dtype=np.int32)
    e.step(a), o.step(a)
    assert o.state[0, 13] == 254  # checked against the unmodified reference (x 248 -> 254)
    assert np.array_equal(e.export(), o.state)


def test_random_states_one_step(emul):
    """Fuzz: random (clamped-to-range) states, far outside what play reaches, stepped once by both."""
    rng = np.random.default_rng(99)
    n = 60_000
    base = po.OracleVecEnv(n, seed=5)
    base.reset()
    st = base.state.copy()
    for k in (0, 13):
        lo, hi = (32, 184) if k == 0 else (248, 400)
        st[:, k + 0] = rng.integers(lo, hi + 1, n)
        st[:, k + 1] = rng.integers(108, 245, n)
        st[:, k + 2] = rng.integers(-16, 17, n)
        st[:, k + 3] = rng.integers(0, 5, n)
        st[:, k + 4] = rng.integers(0, 5, n)
        st[:, k + 5] = rng.integers(0, 6, n)
        st[:, k + 6] = rng.choice([-1, 1], n)
        st[:, k + 7] = rng.integers(-1, 2, n)
        st[:, k + 8] = rng.integers(-2, 4, n)
        st[:, k + 9] = rng.integers(0, 2, n)
        st[:, k + 10] = rng.integers(0, 5, n)
        st[:, k + 11] = rng.integers(0, 2, n)
        st[:, k + 12] = rng.integers(0, 2, n)
    st[:, 26] = rng.integers(20, 433, n)
    st[:, 27] = rng.integers(-200, 253, n)
    st[:, 28] = rng.integers(-20, 21, n)
    st[:, 29] = rng.integers(-300, 301, n)
    st[:, 30:34] = rng.integers(0, 253, (n, 4))
    st[:, 34] = rng.integers(0, 2, n)
    st[:, 35] = rng.integers(0, 433, n)
    st[:, 36] = rng.integers(0, 433, n)
    st[:, 37:39] = rng.integers(0, 4, (n, 2))
    st[:, 39] = rng.integers(0, 2, n) * (rng.integers(0, 4, n) == 0)
    st[:, 41] = rng.integers(0, 2, n)
    for name in ("ai_vs_ai", "ai_p2_random_serve", "random18"):
        cfg = CONFIGS[name]
        e = EmulVecEnv(emul, n, 0, **cfg)
        o = po.OracleVecEnv(n, seed=0, **cfg)
        e.load(st)
        o.state[:] = st
        assert np.array_equal(e.export(), st)  # pack/unpack round trip of every field
        for t in range(3):
            a = synth_actions_numpy(4, 0, n, t, 18)
            obs, base_r, done = e.step(a)
            o_obs, o_rew, o_done = o.step(a)
            bad = np.nonzero((obs != o_obs).any(axis=(1, 2)))[0]
            assert len(bad) == 0, (name, t, st[bad[0]].tolist())
            assert np.array_equal(e.export(), o.state)


def test_rng_rejection_corners(emul):
    """The fused draw sites (two boldness draws; `integers(0, 20)` then `integers(0, 2)`) leave through the
    general path when Lemire's rejection test applies to a value they look at. Force it: the buffered half
    is set to the values v whose v * HIGH mod 2**32 < HIGH (and their neighbours) for every bound in use,
    on states that are about to start a round or stand within reach of their target."""
    rng = np.random.default_rng(7)
    corners = sorted({(k * 2**32 + h - 1) // h + d for h in (2, 3, 5, 20) for k in range(h) for d in (-1, 0, 1)}
                     - {-1, 2**32})
    corners = np.array([c % 2**32 for c in corners], dtype=np.uint64)
    n = 4096
    for name in ("ai_vs_ai", "ai_p2_random_serve", "random18"):
        cfg = CONFIGS[name]
        o = po.OracleVecEnv(n, seed=11, **cfg)
        o.reset()
        for t in range(40):  # into play: some envs near a round end, computer players at their targets
            o.step(synth_actions_numpy(3, 0, n, t, 18))
        st = o.state.copy()
        st[:, 39] = rng.integers(0, 2, n)  # round_ended: half of the envs draw their new-round values next
        st[:, 39] *= 1 - st[:, 40]
        st[:, 50] = 1
        st[:, 51] = corners[rng.integers(0, len(corners), n)].astype(np.uint32).view(np.int32)
        e = EmulVecEnv(emul, n, 0, **cfg)
        e.load(st)
        o.state[:] = st
        for t in range(6):
            a = synth_actions_numpy(5, 0, n, t, 18)
            obs, base_r, done = e.step(a)
            o_obs, o_rew, o_done = o.step(a)
            assert np.array_equal(obs, o_obs), (name, t)
            assert np.array_equal(e.export(), o.state), (name, t)


def test_normalisation_is_correctly_rounded_for_every_representable_value(emul):
    """The division-free float32 NormalizeObservation (reciprocal multiply + two FMAs) equals
    float32(float64 division) — what `.astype(float32)` of the reference wrapper's output gives — for
    every value a packed field can hold (the whole int16 range for each of the 35 elements)."""
    emul.emul_normalize.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 3 + [ctypes.c_int]
    vals = np.arange(-32768, 32768, dtype=np.int32)
    u = np.repeat(vals[:, None], 35, axis=1).copy()
    f32 = np.zeros(u.shape, dtype=np.float32)
    f64 = np.zeros(u.shape, dtype=np.float64)
    emul.emul_normalize(len(u), _p(u), _p(f32), _p(f64), 1)
    low = np.array(po.convert_obs(np.zeros((1, 2, 35), np.int32), np.float64, True)[0, 0])  # = -low / range
    ref64 = po.normalize_obs(np.concatenate([u, u], axis=1).reshape(-1, 2, 35))[:, 0, :]
    assert low.shape == (35,)
    assert np.array_equal(f64.view(np.uint64), ref64.view(np.uint64))
    assert np.array_equal(f32.view(np.uint32), ref64.astype(np.float32).view(np.uint32))
    emul.emul_normalize(len(u), _p(u), _p(f32), _p(f64), 0)
    assert np.array_equal(f32, u.astype(np.float32)) and np.array_equal(f64, u.astype(np.float64))


def test_synth_actions_match(emul):
    for env, frame, agent in [(0, 0, 0), (5, 77, 1), (2**40 + 3, 2**33, 0)]:
        for n_actions in (13, 18):
            assert emul.emul_synth_action(91, env, frame, agent, n_actions) == po.synth_action(91, env, frame, agent,
                                                                                               n_actions)


# ---- the policy kernels' samplers (csrc/pz_policy.cuh) -------------------------------------------------------
@pytest.mark.parametrize("n_actions,generic", [(18, 0), (18, 1), (13, 1), (7, 1), (24, 1), (1, 1)])
def test_policy_inverse_cdf_sampler_matches_numpy_restatement(emul, n_actions, generic):
    """sample_inverse_cdf — the code the tcgen05 policy kernel's threads run — on random logits against
    policy.inverse_cdf_reference: counters (incl. 64-bit seeds / steps and the global env offset), both template
    forms (18 candidates; 24 with a run-time count), peaked and flat distributions, -inf logits. libm's exp2f stands
    in for MUFU.EX2 here and numpy's exp2 there, so a sample may differ where the target falls within rounding of a
    boundary of the cumulative sums: next to never, and then by one action."""
    from pikazoo_b200.policy import inverse_cdf_reference

    emul.emul_policy_inverse_cdf.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p]
    n = 60_000
    rng = np.random.default_rng(n_actions + generic)
    logits = (rng.normal(size=(n, 2, n_actions)) * rng.choice([0.1, 1.0, 6.0], size=(n, 1, 1))).astype(np.float32)
    logits[rng.random(logits.shape) < 0.02] = -np.inf
    logits[:, :, 0] = np.where(np.isinf(logits).all(axis=-1), 0.0, logits[:, :, 0])  # at least one finite logit
    for seed, step, first in ((0, 0, 0), (2**63 + 11, 2**40 + 3, 10**9)):
        got = np.zeros((n, 2), dtype=np.int32)
        emul.emul_policy_inverse_cdf(n, n_actions, _p(logits), seed, step, first, generic, _p(got))
        want = inverse_cdf_reference(logits, seed, step, first)
        assert got.min() >= 0 and got.max() < n_actions
        bad = np.argwhere(got != want)
        assert len(bad) <= 4, len(bad)
        for e, a in bad:
            assert abs(int(got[e, a]) - int(want[e, a])) == 1
        assert not np.isinf(np.take_along_axis(logits, got[..., None].astype(np.int64), axis=-1)).any()  # p = 0 is never drawn
    if n_actions > 1:  # and the frequencies are softmax(logits)
        flat = np.broadcast_to(rng.normal(size=(1, 1, n_actions)).astype(np.float32), (n, 2, n_actions)).copy()
        got = np.zeros((n, 2), dtype=np.int32)
        emul.emul_policy_inverse_cdf(n, n_actions, _p(flat), 5, 6, 7, generic, _p(got))
        p = np.exp(flat[0, 0].astype(np.float64))
        p /= p.sum()
        freq = np.bincount(got.ravel(), minlength=n_actions) / got.size
        assert np.all(np.abs(freq - p) < 5 * np.sqrt(p * (1 - p) / got.size) + 1e-6)


def test_policy_gumbel_keys_match_numpy_restatement(emul):
    """gumbel_key — the mma.sync policy kernel's noise — against policy.gumbel_noise_reference (libm log2f here,
    numpy's log2 there: a few float32 ulp)."""
    from pikazoo_b200.policy import gumbel_noise_reference

    emul.emul_policy_gumbel_keys.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64,
                                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    n = 20_000
    logits = np.random.default_rng(3).normal(size=(n, 2, 18)).astype(np.float32)
    for seed, step, first in ((1, 2, 3), (2**64 - 1, 2**50, 123456)):
        keys = np.zeros_like(logits)
        emul.emul_policy_gumbel_keys(n, 18, _p(logits), seed, step, first, _p(keys))
        want = logits + gumbel_noise_reference(seed, step, first, n, 18)
        assert np.abs(keys - want).max() < 2e-5


def test_one_step_decode_equals_key_table_decode(emul):
    """decode_input (five per-key action masks: what the step and rollout kernels run) == get_input(decode_keys(.))
    (the 5-bit key table of action_key_map, pikazoo_env.py:119-141 + PikaUserInput.get_input, physics.py:59-99) for
    every action in and out of range, both agents, with and without SimplifyAction, both previous key states."""
    out = np.zeros(10, dtype=np.int32)
    for agent in (0, 1):
        for simplify in (0, 1):
            for action in list(range(-3, 40)) + [2**31 - 1, -2**31, 255, 1000]:
                for keyprev in (0, 1):
                    emul.emul_decode_both(agent, simplify, ctypes.c_int(action), keyprev, _p(out))
                    assert np.array_equal(out[:5], out[5:]), (agent, simplify, action, keyprev, out)
                    n = 13 if simplify else 18
                    assert out[4] == (0 if 0 <= action < n else 1)


def test_player_move_table_forms_equal_arithmetic(emul):
    """player_move reading the sprite animation (physics.py:524-552) from a table — the one the K-frame kernels fill in
    shared memory (anim_fill) and the compile-time one of the per-step kernels (g_anim_table) — == player_move computing
    it, for every (state, frame, delay, arm) x positions x velocities x lying / diving fields x all 18 inputs, both
    players' court halves."""
    emul.emul_player_move_forms.restype = ctypes.c_int64
    bad = ctypes.c_int64(-1)
    checked = emul.emul_player_move_forms(ctypes.byref(bad))
    assert checked > 5_000_000 and bad.value == 0, (checked, bad.value)
