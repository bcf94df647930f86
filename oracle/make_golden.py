"""TEST INFRASTRUCTURE — generates tests/golden/sessions.json from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # full set, 1,000+ games, ~5 min on 8 cores
    python -m oracle.make_golden --quick    # a few sessions, for a smoke check

A *session* is one reference env object, seeded by protocol S0 (oracle/ref_harness.py),
played for ``episodes`` consecutive games (``reset()`` on the SAME object between games, so
the reference's carry-over across reset() is exercised) with a deterministic action stream
(``pk_synth_action``, integer-only, restated in oracle/synth.py). A *group* is a set of
sessions sharing one config whose seeds are base_seed + i, i.e. exactly one batched product
run with ``num_envs = n`` and ``seed = base_seed``.

For every session the file stores the per-episode (frames, score) and a sha256 over

    reset obs (2x35 int32 LE) | per step: obs (2x35 int32), rewards (2 float64), terminated byte
    | after a terminated step: the reset obs of the next episode (auto-reset call)

While generating, the C oracle (oracle/pika_oracle.c) is stepped alongside the reference
and compared on EVERY frame: observations, rewards, termination and the full 52-word hidden
state including the PCG64 stream. A mismatch aborts. The output therefore records
``oracle_checked: true`` for each group.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import pyoracle as po  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle.synth import synth_action  # noqa: E402

SHAPED = ((0.1, 0.2, 0.3, 0.4, -0.1, -0.2, -0.3, -0.4), 216, 176)

# name, config kwargs, action mode ("noop" | "synth"), base_seed, n sessions, episodes each, max calls
GROUPS_FULL = [
    ("ai_vs_ai_ws15_winner", dict(is_player1_computer=True, is_player2_computer=True, winning_score=15, serve="winner"),
     "noop", 0, 100, 1, 30000),
    ("random_ws15_winner", dict(winning_score=15, serve="winner"), "synth", 1000, 300, 1, 30000),
    ("simplify_shaped_ws15", dict(winning_score=15, serve="winner", simplify_action=True,
                                  reward_by_ball_position=SHAPED), "synth", 2000, 200, 1, 30000),
    ("random_ws5_serve_random", dict(winning_score=5, serve="random"), "synth", 3000, 40, 5, 30000),
    ("ai_p1_vs_random_ws7_alternate", dict(is_player1_computer=True, winning_score=7, serve="alternate"),
     "synth", 4000, 100, 1, 30000),
    ("random_vs_ai_p2_ws7_random", dict(is_player2_computer=True, winning_score=7, serve="random"),
     "synth", 5000, 100, 1, 30000),
    ("ai_vs_ai_ws3_random_multi", dict(is_player1_computer=True, is_player2_computer=True, winning_score=3,
                                       serve="random"), "noop", 6000, 10, 4, 30000),
]

GROUPS_QUICK = [
    ("ai_vs_ai_ws15_winner", GROUPS_FULL[0][1], "noop", 0, 2, 1, 30000),
    ("random_ws15_winner", GROUPS_FULL[1][1], "synth", 1000, 4, 1, 30000),
    ("simplify_shaped_ws15", GROUPS_FULL[2][1], "synth", 2000, 4, 1, 30000),
    ("random_ws5_serve_random", GROUPS_FULL[3][1], "synth", 3000, 2, 3, 30000),
]

ACTION_SEED = 0x5EED


def _hash_obs(h, obs):
    h.update(np.asarray(obs["player_1"]).astype("<i4").tobytes())
    h.update(np.asarray(obs["player_2"]).astype("<i4").tobytes())


def run_session(args):
    """One reference env, checked frame by frame against the C oracle."""
    cfg, action_mode, seed, env_index, episodes, max_calls = args
    env = rh.make_env(seed, **cfg)
    raw = env.raw
    n_actions = 13 if cfg.get("simplify_action") else 18

    ocfg = po.make_config(**cfg)
    ostate = np.zeros(po.ENV_WORDS, dtype=np.int32)
    oobs = np.zeros(70, dtype=np.int32)
    orew = np.zeros(2, dtype=np.float64)
    oterm = np.zeros(1, dtype=np.uint8)
    L = po.lib()
    L.pk_init(po._p(ostate), seed)

    def check(obs, rew, term, where):
        ref_obs = np.concatenate([obs["player_1"], obs["player_2"]]).astype(np.int32)
        if not np.array_equal(ref_obs, oobs):
            raise AssertionError(f"obs mismatch {where}: {np.nonzero(ref_obs != oobs)[0]}")
        ref_state = rh.unpacked_state(raw)
        if not np.array_equal(ref_state, ostate[:52]):
            raise AssertionError(f"state mismatch {where}: words {np.nonzero(ref_state != ostate[:52])[0]}")
        if rew is not None:
            r = np.array([rew["player_1"], rew["player_2"]], dtype=np.float64)
            if not np.array_equal(r, orew) or bool(term) != bool(oterm[0]):
                raise AssertionError(f"reward/term mismatch {where}: {r} {orew} {term} {oterm}")

    h = hashlib.sha256()
    import ctypes

    cref = ctypes.byref(ocfg)
    obs, _ = env.reset()
    L.pk_reset(po._p(ostate), cref, po._p(oobs))
    check(obs, None, None, (seed, "reset0"))
    _hash_obs(h, obs)

    results = []
    calls = 0
    frame = 0  # global call counter drives the action stream (resets consume an index too)
    ep_frames = 0
    while len(results) < episodes and calls < max_calls:
        if action_mode == "noop":
            a1 = a2 = 0
        else:
            a1 = synth_action(ACTION_SEED, env_index, frame, 0, n_actions)
            a2 = synth_action(ACTION_SEED, env_index, frame, 1, n_actions)
        obs, rew, terms, _, _ = env.step({"player_1": a1, "player_2": a2})
        rc = L.pk_step(po._p(ostate), cref, a1, a2, po._p(oobs), po._p(orew), po._p(oterm))
        assert rc == 0
        calls += 1
        frame += 1
        ep_frames += 1
        term = bool(terms["player_1"])
        check(obs, rew, term, (seed, len(results), ep_frames))
        _hash_obs(h, obs)
        h.update(np.array([rew["player_1"], rew["player_2"]], dtype="<f8").tobytes())
        h.update(bytes([int(term)]))
        if term:
            results.append({"frames": ep_frames, "scores": [int(raw.scores[0]), int(raw.scores[1])]})
            ep_frames = 0
            if len(results) < episodes and calls < max_calls:
                obs, _ = env.reset()  # the product's auto-reset call
                L.pk_reset(po._p(ostate), cref, po._p(oobs))
                calls += 1
                frame += 1
                check(obs, None, None, (seed, len(results), "reset"))
                _hash_obs(h, obs)
    return {
        "seed": seed,
        "episodes": results,
        "calls": calls,
        "unfinished_frames": ep_frames,
        "sha256": h.hexdigest(),
        "final_state": [int(v) for v in ostate[:52]],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(_ROOT, "tests", "golden", "sessions.json"))
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    if not rh.reference_available():
        raise SystemExit("reference not present; golden fixtures can only be generated in the build container")
    po.build()
    groups = GROUPS_QUICK if a.quick else GROUPS_FULL
    out = {
        "generator": "oracle/make_golden.py",
        "reference": "helpingstar/pika-zoo @ /root/reference (unmodified), numpy " + np.__version__,
        "seeding": "S0: env.np_random.bit_generator.state = np.random.PCG64(seed).state; reset()",
        "action_seed": ACTION_SEED,
        "hash": "sha256(reset obs | per step: obs1, obs2 int32 LE, rewards 2xfloat64 LE, terminated byte | reset obs after a terminated step)",
        "groups": [],
    }
    total_games = 0
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        for name, cfg, mode, base_seed, n, episodes, max_calls in groups:
            jobs = [(cfg, mode, base_seed + i, i, episodes, max_calls) for i in range(n)]
            sessions = pool.map(run_session, jobs, chunksize=1)
            games = sum(len(s["episodes"]) for s in sessions)
            total_games += games
            jcfg = dict(cfg)
            if "reward_by_ball_position" in jcfg:
                add, xl, yl = jcfg["reward_by_ball_position"]
                jcfg["reward_by_ball_position"] = [list(add), xl, yl]
            out["groups"].append({
                "name": name, "config": jcfg, "action_mode": mode, "base_seed": base_seed,
                "num_envs": n, "episodes_per_env": episodes, "max_calls": max_calls,
                "oracle_checked": True, "games": games, "sessions": sessions,
            })
            print(f"{name}: {n} sessions, {games} games, {time.time() - t0:.1f}s", flush=True)
    out["total_games"] = total_games
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {a.out}: {total_games} games, {os.path.getsize(a.out) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
