"""Shared test machinery: replay golden sessions through any batched stepper and hash the
per-env streams exactly as oracle/make_golden.py did with the reference."""

from __future__ import annotations

import hashlib

import numpy as np

from oracle.synth import synth_actions_numpy


def group_kwargs(group):
    cfg = dict(group["config"])
    if "reward_by_ball_position" in cfg:
        add, xl, yl = cfg["reward_by_ball_position"]
        cfg["reward_by_ball_position"] = (tuple(add), xl, yl)
    return cfg


class SessionRecorder:
    """Per-env sha256 / episode bookkeeping for a batch stepped in lock-step with auto-reset."""

    def __init__(self, n, episodes_per_env, max_calls, normalized=False, record=False):
        self.n = n
        self.obs_dt = "<f8" if normalized else "<i4"   # NormalizeObservation yields float64
        self.record = record                             # RecordEpisodeStatistics: hash r / l on termination
        self.h = [hashlib.sha256() for _ in range(n)]
        self.episodes = [[] for _ in range(n)]
        self.calls = np.zeros(n, dtype=np.int64)
        self.ep_frames = np.zeros(n, dtype=np.int64)
        self.pending_reset = np.zeros(n, dtype=bool)  # this env's next call is its auto-reset
        self.active = np.ones(n, dtype=bool)
        self.episodes_per_env = episodes_per_env
        self.max_calls = max_calls

    def on_reset(self, obs):
        assert np.asarray(obs).dtype == np.dtype(self.obs_dt), np.asarray(obs).dtype
        obs = np.ascontiguousarray(obs, dtype=self.obs_dt)
        for i in range(self.n):
            self.h[i].update(obs[i].tobytes())

    def on_step(self, obs, reward, done, scores_fn, episode_fn=None):
        """obs [n,2,35] int32 (float64 when normalised), reward [n,2] float64, done [n] bool — outputs of
        one batched call; episode_fn() -> (returns [n,2] float64, lengths [n]) when recording."""
        assert np.asarray(obs).dtype == np.dtype(self.obs_dt), np.asarray(obs).dtype
        obs = np.ascontiguousarray(obs, dtype=self.obs_dt)
        episode = None
        reward = np.ascontiguousarray(reward, dtype="<f8")
        scores = None
        for i in np.nonzero(self.active)[0]:
            self.calls[i] += 1
            if self.pending_reset[i]:
                self.h[i].update(obs[i].tobytes())
                self.pending_reset[i] = False
            else:
                self.ep_frames[i] += 1
                self.h[i].update(obs[i].tobytes())
                self.h[i].update(reward[i].tobytes())
                self.h[i].update(bytes([int(done[i])]))
                if done[i]:
                    if scores is None:
                        scores = scores_fn()
                    ep = {"frames": int(self.ep_frames[i]), "scores": [int(scores[i][0]), int(scores[i][1])]}
                    if self.record:
                        if episode is None:
                            episode = episode_fn()
                        r = np.ascontiguousarray(episode[0][i], dtype="<f8")
                        self.h[i].update(r.tobytes())
                        self.h[i].update(np.array([episode[1][i]], dtype="<i4").tobytes())
                        ep["returns"] = [float(r[0]), float(r[1])]
                    self.episodes[i].append(ep)
                    self.ep_frames[i] = 0
                    self.pending_reset[i] = True
            if self.calls[i] >= self.max_calls:
                self.active[i] = False
            elif len(self.episodes[i]) >= self.episodes_per_env:
                # the golden generator stops right after the terminating step
                self.active[i] = False

    def all_done(self):
        return not self.active.any()


def replay_group(group, make_stepper, check_final_state=True):
    """make_stepper(n, base_seed, cfg_kwargs) -> object with
         reset() -> obs ; step(actions int32 [n,2]) -> (obs, reward f64, done) ; scores() -> [n,2];
         final_state() -> [n, >=52] int32 (optional)
    Returns a list of mismatch descriptions (empty = parity)."""
    n = group["num_envs"]
    cfg = group_kwargs(group)
    n_actions = 13 if cfg.get("simplify_action") else 18
    stepper = make_stepper(n, group["base_seed"], cfg)
    rec = SessionRecorder(n, group["episodes_per_env"], group["max_calls"],
                          normalized=bool(cfg.get("normalize_observation")),
                          record=bool(cfg.get("record_episode_statistics")))
    rec.on_reset(stepper.reset())
    frame = 0
    final_states = [None] * n
    while not rec.all_done():
        if group["action_mode"] == "noop":
            actions = np.zeros((n, 2), dtype=np.int32)
        else:
            actions = synth_actions_numpy(0x5EED, 0, n, frame, n_actions)
        was_active = rec.active.copy()
        obs, reward, done = stepper.step(actions)
        rec.on_step(obs, reward, done, stepper.scores, getattr(stepper, "episode", None))
        frame += 1
        finished_now = was_active & ~rec.active
        if check_final_state and finished_now.any():
            st = stepper.final_state()
            for i in np.nonzero(finished_now)[0]:
                final_states[i] = np.array(st[i][:52], dtype=np.int32)
    bad = []
    for i, sess in enumerate(group["sessions"]):
        if rec.episodes[i] != sess["episodes"]:
            bad.append(f"{group['name']} env {i}: episodes {rec.episodes[i]} != {sess['episodes']}")
        elif rec.h[i].hexdigest() != sess["sha256"]:
            bad.append(f"{group['name']} env {i}: sha256 differs")
        elif check_final_state and final_states[i] is not None:
            ref = np.array(sess["final_state"], dtype=np.int64).astype(np.uint32).view(np.int32)
            if not np.array_equal(final_states[i], ref):
                bad.append(f"{group['name']} env {i}: final state words {np.nonzero(final_states[i] != ref)[0]}")
    return bad


class OracleStepper:
    def __init__(self, n, base_seed, cfg):
        from oracle import pyoracle as po

        self.env = po.OracleVecEnv(n, seed=base_seed, autoreset=True, **cfg)
        self.normalized = bool(cfg.get("normalize_observation"))

    def _obs(self):
        return self.env.normalized_obs() if self.normalized else self.env.obs

    def reset(self):
        self.env.reset()
        return self._obs()

    def step(self, actions):
        _, reward, done = self.env.step(actions)
        return self._obs(), reward, done

    def episode(self):
        return self.env.episode_return, self.env.episode_length

    def scores(self):
        return self.env.state[:, 37:39]

    def final_state(self):
        return self.env.state


class CudaStepper:
    landing_tables = "auto"

    def __init__(self, n, base_seed, cfg):
        import torch
        import pikazoo_b200

        self.torch = torch
        self.env = pikazoo_b200.PikaVecEnv(n, device="cuda", seed=base_seed, autoreset=True,
                                           reward_dtype=torch.float64, landing_tables=self.landing_tables, **cfg)

    def reset(self):
        return self.env.reset().cpu().numpy()

    def step(self, actions):
        a = self.torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int32)).to(self.env.device)
        obs, reward, done = self.env.step(a)
        return obs.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy()

    def scores(self):
        return self.env.scores().cpu().numpy()

    def final_state(self):
        return self.env.export_state().cpu().numpy()


class CudaStepperTables(CudaStepper):
    """Computer players read the memoised trajectory tables."""

    landing_tables = True


class CudaStepperIterative(CudaStepper):
    """Computer players iterate every trajectory simulation (PZ_FLAG_NO_TABLES)."""

    landing_tables = False


def survey_known_answers(make_env, n, cfg, actions_for, max_frames=25_000):
    """SURVEY.md §8(c) known answers: env i is seeded i (protocol S0); hash = sha256 over obs1, obs2
    (int32) after reset, then per step obs1, obs2, bytes([reward_p1 + 1]); first 16 hex digits.
    make_env(n, seed, **cfg) -> object with reset() -> obs [n,2,35], step(a) -> (obs, reward, done) as
    numpy, scores() -> [n,2]. actions_for(i, t) -> (a1, a2). Returns [(frames, [s1, s2], hash16) | None]."""
    env = make_env(n, 0, cfg)
    obs = np.asarray(env.reset())
    hs = [hashlib.sha256() for _ in range(n)]
    for i in range(n):
        hs[i].update(np.ascontiguousarray(obs[i], dtype="<i4").tobytes())
    out = [None] * n
    t = 0
    while any(o is None for o in out) and t < max_frames:
        a = np.array([actions_for(i, t) for i in range(n)], dtype=np.int32)
        obs, rew, done = env.step(a)
        obs, rew, done = np.asarray(obs), np.asarray(rew), np.asarray(done)
        t += 1
        scores = None
        for i in range(n):
            if out[i] is not None:
                continue
            hs[i].update(np.ascontiguousarray(obs[i], dtype="<i4").tobytes())
            hs[i].update(bytes([int(rew[i][0]) + 1]))
            if done[i]:
                if scores is None:
                    scores = np.asarray(env.scores())
                out[i] = (t, [int(scores[i][0]), int(scores[i][1])], hs[i].hexdigest()[:16])
    return out


class _PerEnvRng:
    """actions of the survey's random-vs-random answers: env i draws
    np.random.default_rng(i + 1000).integers(0, 18, size=2) once per step"""

    def __init__(self, n):
        self.rngs = [np.random.default_rng(i + 1000) for i in range(n)]
        self.cache = {}

    def __call__(self, i, t):
        if (i, t) not in self.cache:
            self.cache[(i, t)] = tuple(int(v) for v in self.rngs[i].integers(0, 18, size=2))
        return self.cache[(i, t)]


SURVEY_AI = {0: (13987, [15, 5], "7c7cc240a767c583"), 1: (11383, [15, 4], "13bf71668641d8b4"), 2: None,
             3: (17039, [15, 2], None)}
SURVEY_RANDOM_WS15 = {0: (1124, [9, 15]), 1: (1075, [13, 15]), 3: (1240, [15, 11])}
SURVEY_RANDOM_WS5 = {0: (445, [5, 4]), 1: (197, [1, 5]), 3: (259, [0, 5])}


def check_survey_known_answers(make_env):
    """The answers SURVEY.md §8(c) recorded from the unmodified reference, reproduced by `make_env`."""
    ai = survey_known_answers(make_env, 4, dict(is_player1_computer=True, is_player2_computer=True, winning_score=15,
                                                serve="winner"), lambda i, t: (0, 0), max_frames=17_100)
    for i, exp in SURVEY_AI.items():
        if exp is None:
            assert ai[i] is None  # seed 2 never terminates
        else:
            assert ai[i][:2] == exp[:2], (i, ai[i])
            assert exp[2] is None or ai[i][2] == exp[2], (i, ai[i])
    for cfg, table in ((dict(), SURVEY_RANDOM_WS15), (dict(winning_score=5, serve="random"), SURVEY_RANDOM_WS5)):
        got = survey_known_answers(make_env, 4, cfg, _PerEnvRng(4), max_frames=3_000)
        for i, exp in table.items():
            assert got[i][:2] == exp, (cfg, i, got[i])
