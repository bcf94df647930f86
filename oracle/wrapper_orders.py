"""TEST INFRASTRUCTURE — wrapper stacks in orders the fused kernel options do not cover: the same
session protocol driven over the unmodified reference (oracle/make_golden.py --orders, build container) and over the
product's facade (tests/test_gpu_wrappers.py), so that both sides hash exactly the same things.

A stack is a list of (kind, *args), INNERMOST first; kinds are the reference's wrapper classes
(pikazoo/wrappers/*.py): simplify, rbbp (RewardByBallPosition), rins (RewardInNormalState), normalize, record.
"""

from __future__ import annotations

import hashlib

import numpy as np

ADD = (0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4)
ADD2 = (1, 2, 3, 4, 5, 6, 7, 8)

STACKS = {
    # RecordEpisodeStatistics records what is BELOW it: unshaped rewards here
    "record_inside_rins": [("record",), ("rins", 0.01)],
    "record_inside_rbbp": [("record",), ("rbbp", ADD)],
    # RewardByBallPosition reads obs[26], obs[27] of the env below it: normalised values, zone always 0
    "normalize_inside_rbbp": [("normalize",), ("rbbp", ADD)],
    # two RewardByBallPosition add up
    "rbbp_twice": [("rbbp", ADD), ("rbbp", ADD2, 100, 100)],
    "rins_twice": [("rins", 0.5), ("rins", 0.25)],
    "normalize_twice": [("normalize",), ("normalize",)],
    "mixed": [("simplify",), ("rbbp", ADD), ("record",), ("normalize",), ("rins", -0.01)],
    # and one the kernel does fuse entirely, for contrast
    "canonical": [("simplify",), ("rins", 0.02), ("rbbp", ADD), ("normalize",), ("record",)],
}

ENV_KW = dict(winning_score=2, serve="random")
SEEDS = (11, 12, 13)
EPISODES = 2
ACTION_SEED = 424242


def build_stack(env, W, stack):
    cls = {"simplify": W.SimplifyAction, "rbbp": W.RewardByBallPosition, "rins": W.RewardInNormalState,
           "normalize": W.NormalizeObservation, "record": W.RecordEpisodeStatistics}
    for kind, *args in stack:
        env = cls[kind](env, *args)
    return env


def run_session(env, n_actions, action_fn, episodes=EPISODES, max_calls=20_000):
    """Plays `episodes` games on `env` (already seeded). action_fn(frame, agent_index) -> action in [0, n_actions).
    Returns per-episode records and a sha256 over everything the wrappers return."""
    h = hashlib.sha256()

    def feed_obs(obs):
        for a in ("player_1", "player_2"):
            h.update(np.asarray(obs[a]).astype("<f8").tobytes())

    obs, _ = env.reset()
    feed_obs(obs)
    out, frame, ep_frames = [], 0, 0
    while len(out) < episodes and frame < max_calls:
        acts = {"player_1": action_fn(frame, 0), "player_2": action_fn(frame, 1)}
        obs, rew, term, trunc, infos = env.step(acts)
        frame += 1
        ep_frames += 1
        feed_obs(obs)
        h.update(np.array([rew["player_1"], rew["player_2"]], dtype="<f8").tobytes())
        h.update(bytes([int(bool(term["player_1"])), int(bool(trunc["player_1"]))]))
        if term["player_1"]:
            ep = {"frames": ep_frames, "score": [int(s) for s in infos["player_1"]["score"]]}
            if "episode" in infos["player_1"]:
                ep["returns"] = [float(infos[a]["episode"]["r"]) for a in ("player_1", "player_2")]
                ep["lengths"] = [int(infos[a]["episode"]["l"]) for a in ("player_1", "player_2")]
                h.update(np.array(ep["returns"], dtype="<f8").tobytes())
            out.append(ep)
            ep_frames = 0
            if len(out) < episodes:
                obs, _ = env.reset()
                feed_obs(obs)
    return {"episodes": out, "calls": frame, "sha256": h.hexdigest()}
