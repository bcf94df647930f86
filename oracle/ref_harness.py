"""TEST INFRASTRUCTURE — never imported by the product path.

Runs the UNMODIFIED reference (helpingstar/pika-zoo, mounted read-only at
/root/reference) in this container so that

  * the C restatement in ``oracle/pika_oracle.c`` can be validated against it, and
  * golden fixtures under ``tests/golden/`` can be generated (``oracle/make_golden.py``).

``gymnasium``, ``pettingzoo`` and ``pygame`` are not installed here and there is no
network, so ~40 lines of stand-ins are injected into ``sys.modules``.
Only what the reference touches with ``render_mode=None`` is provided:

  gymnasium.spaces.{Discrete,Box,Space}     pikazoo_env.py:4,90-95,484-565
  gymnasium.utils.seeding.np_random         pikazoo_env.py:570-571
  gymnasium.logger                          pikazoo_env.py:366 (render only)
  pettingzoo.ParallelEnv                    pikazoo_env.py:6,72
  pettingzoo.utils.BaseParallelWrapper      wrappers/*.py
  pettingzoo.utils.env.ParallelEnv          wrappers/*.py
  pygame                                    pikazoo_env.py:21 (import only)

``/root/reference`` does not exist on the GPU box; there the harness falls back to the staged copy of
the reference's Python sources under ``oracle/_ref/`` (``oracle/stage_ref.py``), which only
``oracle/time_reference.py`` (bench.py's CPU legs) uses. Tests that need the reference are skipped
when neither is present.

Seeding protocol S0 (the only way to seed the unmodified reference, whose
``reset(seed=...)`` ignores its argument, pikazoo_env.py:149): construct, then
``env.np_random.bit_generator.state = np.random.PCG64(seed).state``, then ``reset()``.
"""

from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """PIKA_REFERENCE_ROOT, else the read-only mount of the build container, else the byte-for-byte copy of
    the reference's .py files that oracle/stage_ref.py stages under oracle/_ref/ (git-ignored; it travels to
    the GPU box with the tree, where /root/reference does not exist)."""
    env = os.environ.get("PIKA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/pikazoo"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pikazoo"))


def _install_stubs() -> None:
    if "pettingzoo" in sys.modules and "gymnasium" in sys.modules:
        return

    # ---- gymnasium -----------------------------------------------------------------
    gymnasium = types.ModuleType("gymnasium")
    spaces = types.ModuleType("gymnasium.spaces")
    utils = types.ModuleType("gymnasium.utils")
    seeding = types.ModuleType("gymnasium.utils.seeding")
    logger = types.ModuleType("gymnasium.logger")

    class Space:
        pass

    class Discrete(Space):
        def __init__(self, n):
            self.n = int(n)
            self._rng = np.random.default_rng()

        def sample(self):
            return int(self._rng.integers(0, self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.asarray(low)
            self.high = np.asarray(high)
            self.shape = tuple(shape) if shape is not None else self.low.shape
            self.dtype = np.dtype(dtype)

    def np_random(seed=None):
        ss = np.random.SeedSequence(seed)
        return np.random.Generator(np.random.PCG64(ss)), ss.entropy

    spaces.Space, spaces.Discrete, spaces.Box = Space, Discrete, Box
    seeding.np_random = np_random
    logger.warn = lambda *a, **k: None
    utils.seeding = seeding
    gymnasium.spaces, gymnasium.utils, gymnasium.logger = spaces, utils, logger

    # ---- pettingzoo ----------------------------------------------------------------
    pettingzoo = types.ModuleType("pettingzoo")
    pz_utils = types.ModuleType("pettingzoo.utils")
    pz_env = types.ModuleType("pettingzoo.utils.env")

    class ParallelEnv:
        pass

    class BaseParallelWrapper(ParallelEnv):
        def __init__(self, env):
            self.env = env
            self.metadata = getattr(env, "metadata", {})
            self.possible_agents = env.possible_agents
            self.agents = env.agents

        def reset(self, seed=None, options=None):
            res = self.env.reset(seed=seed, options=options)
            self.agents = self.env.agents
            return res

        def step(self, actions):
            res = self.env.step(actions)
            self.agents = self.env.agents
            return res

        def observation_space(self, agent):
            return self.env.observation_space(agent)

        def action_space(self, agent):
            return self.env.action_space(agent)

        @property
        def unwrapped(self):
            e = self.env
            while hasattr(e, "env"):
                e = e.env
            return e

    pettingzoo.ParallelEnv = ParallelEnv
    pz_utils.BaseParallelWrapper = BaseParallelWrapper
    pz_env.ParallelEnv = ParallelEnv
    pz_utils.env = pz_env
    pettingzoo.utils = pz_utils

    # pygame: the recording stand-in of oracle/pygame_stub.py (an empty module would do for render_mode=None; with a
    # render mode the reference's draw() runs against it and its blits are logged as display lists)
    if "pygame" not in sys.modules:
        from oracle import pygame_stub

        pygame_stub.install(sys.modules)

    for name, mod in {
        "gymnasium": gymnasium,
        "gymnasium.spaces": spaces,
        "gymnasium.utils": utils,
        "gymnasium.utils.seeding": seeding,
        "gymnasium.logger": logger,
        "pettingzoo": pettingzoo,
        "pettingzoo.utils": pz_utils,
        "pettingzoo.utils.env": pz_env,
    }.items():
        sys.modules.setdefault(name, mod)


def load_reference():
    """Return (pikazoo_v0 module, wrappers module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from pikazoo import pikazoo_v0  # noqa: E402
    import pikazoo.wrappers as wrappers  # noqa: E402

    return pikazoo_v0, wrappers


def make_env(seed, *, simplify_action=False, reward_by_ball_position=None, reward_in_normal_state=None,
             normal_state_first=False, normalize_observation=False, record_episode_statistics=False, **kwargs):
    """Construct a reference env (optionally wrapped) seeded by protocol S0.

    reward_by_ball_position: None or (additional_reward[8], x_line, y_line).
    reward_in_normal_state: None or the reward; normal_state_first puts that wrapper INSIDE
    RewardByBallPosition instead of outside it. Wrapper order, innermost first:
    SimplifyAction, [RewardInNormalState], RewardByBallPosition, [RewardInNormalState],
    NormalizeObservation, RecordEpisodeStatistics — all the reference's own classes.
    Returns the outermost env; ``.unwrapped`` / ``raw`` attribute gives the raw_env.
    """
    pikazoo_v0, wrappers = load_reference()
    raw = pikazoo_v0.env(**kwargs)
    raw.np_random.bit_generator.state = np.random.PCG64(int(seed)).state
    env = raw
    if simplify_action:
        env = wrappers.SimplifyAction(env)
    if reward_in_normal_state is not None and normal_state_first:
        env = wrappers.RewardInNormalState(env, reward_in_normal_state)
    if reward_by_ball_position is not None:
        add, x_line, y_line = reward_by_ball_position
        env = wrappers.RewardByBallPosition(env, tuple(add), x_line, y_line)
    if reward_in_normal_state is not None and not normal_state_first:
        env = wrappers.RewardInNormalState(env, reward_in_normal_state)
    if normalize_observation:
        env = wrappers.NormalizeObservation(env)
    if record_episode_statistics:
        env = wrappers.RecordEpisodeStatistics(env)
    env.raw = raw
    return env


def unpacked_state(raw) -> np.ndarray:
    """The 52-word parity state of SURVEY.md §8(a) read off a reference raw_env.

    Word order is the one ``oracle/pika_oracle.h`` (struct pk_env) and the CUDA
    library's pz_export_state use.
    """
    ph = raw.physics
    out = []
    for i, p in enumerate((ph.player1, ph.player2)):
        out += [
            p.x, p.y, p.y_velocity, p.state, p.frame_number, p.delay_before_next_frame,
            p.normal_status_arm_swing_direction, p.diving_direction, p.lying_down_duration_left,
            int(p.is_collision_with_ball_happened), int(p.computer_boldness),
            int(p.computer_where_to_stand_by),
            int(raw.keyboard_array[i].power_hit_key_is_down_previous),
        ]
    b = ph.ball
    out += [
        b.x, b.y, b.x_velocity, b.y_velocity, b.previous_x, b.previous_y,
        b.previous_previous_x, b.previous_previous_y, int(b.is_power_hit),
        b.expected_landing_point_x, b.punch_effect_x,
    ]
    out += [raw.scores[0], raw.scores[1], int(raw.round_ended), int(raw.game_ended),
            int(raw.is_player2_serve)]
    st = raw.np_random.bit_generator.state
    s, inc = st["state"]["state"], st["state"]["inc"]
    out += [(s >> (32 * k)) & 0xFFFFFFFF for k in range(4)]
    out += [(inc >> (32 * k)) & 0xFFFFFFFF for k in range(4)]
    out += [int(st["has_uint32"]), int(st["uinteger"])]
    return np.array(out, dtype=np.int64).astype(np.uint32).view(np.int32)


class TrajectoryHasher:
    """sha256 over obs1,obs2 (int32) after reset, then per step obs1, obs2, byte(reward_p1+1).

    Same definition as SURVEY.md §8(c) so the survey's known answers can be reused.
    """

    def __init__(self):
        self.h = hashlib.sha256()

    def reset(self, obs1, obs2):
        self.h.update(np.asarray(obs1).astype(np.int32).tobytes())
        self.h.update(np.asarray(obs2).astype(np.int32).tobytes())

    def step(self, obs1, obs2, reward_p1):
        self.reset(obs1, obs2)
        self.h.update(bytes([int(reward_p1) + 1]))

    def hexdigest16(self):
        return self.h.hexdigest()[:16]


def play_game(seed, actions_fn, max_frames=200_000, record=False, **env_kwargs):
    """Play one game from reset until termination with the reference.

    actions_fn(frame_index) -> (a1, a2). Returns dict(frames, scores, hash16, [trace]).
    trace (record=True) = list of (obs1, obs2, r1, r2, terminated) per step, with the
    reset observation first as (obs1, obs2, 0, 0, False).
    """
    env = make_env(seed, **env_kwargs)
    obs, _ = env.reset()
    hasher = TrajectoryHasher()
    hasher.reset(obs["player_1"], obs["player_2"])
    trace = [(obs["player_1"].copy(), obs["player_2"].copy(), 0, 0, False)] if record else None
    frames = 0
    term = False
    while not term and frames < max_frames:
        a1, a2 = actions_fn(frames)
        obs, rew, terms, _, _ = env.step({"player_1": a1, "player_2": a2})
        frames += 1
        term = terms["player_1"]
        r1 = rew["player_1"]
        if float(r1).is_integer():
            hasher.step(obs["player_1"], obs["player_2"], int(r1))
        else:
            hasher.reset(obs["player_1"], obs["player_2"])
        if record:
            trace.append((obs["player_1"].copy(), obs["player_2"].copy(), r1, rew["player_2"], term))
    return {
        "frames": frames,
        "scores": list(env.raw.scores),
        "terminated": bool(term),
        "hash16": hasher.hexdigest16(),
        "trace": trace,
        "env": env,
    }
