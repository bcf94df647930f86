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

# The other wrappers of the reference, stacked: tests/golden/wrappers.json
W1 = dict(winning_score=5, serve="random", simplify_action=True, reward_by_ball_position=SHAPED,
          reward_in_normal_state=-0.01, normalize_observation=True, record_episode_statistics=True)
W2 = dict(winning_score=5, serve="winner", reward_by_ball_position=((0.5, 0, -0.5, 0.25, 0, 0.125, 0, -1), 200, 150),
          reward_in_normal_state=0.002, normal_state_first=True, record_episode_statistics=True)
W3 = dict(winning_score=3, serve="alternate", is_player2_computer=True, normalize_observation=True,
          record_episode_statistics=True)
W4 = dict(winning_score=2, serve="winner", is_player1_computer=True, is_player2_computer=True,
          reward_in_normal_state=1, normalize_observation=True, record_episode_statistics=True)
GROUPS_WRAPPERS = [
    ("normalize_rins_outer_shaped_simplify_ws5", W1, "synth", 7000, 60, 3, 30000),
    ("rins_inner_shaped_record_ws5", W2, "synth", 8000, 60, 2, 30000),
    ("normalize_record_ai_p2_ws3", W3, "synth", 9000, 40, 2, 30000),
    ("normalize_rins_int_ai_vs_ai_ws2", W4, "noop", 9500, 16, 2, 30000),
]

GROUPS_QUICK = [
    ("ai_vs_ai_ws15_winner", GROUPS_FULL[0][1], "noop", 0, 2, 1, 30000),
    ("random_ws15_winner", GROUPS_FULL[1][1], "synth", 1000, 4, 1, 30000),
    ("simplify_shaped_ws15", GROUPS_FULL[2][1], "synth", 2000, 4, 1, 30000),
    ("random_ws5_serve_random", GROUPS_FULL[3][1], "synth", 3000, 2, 3, 30000),
]

ACTION_SEED = 0x5EED


def _hash_obs(h, obs, normalized=False):
    dt = "<f8" if normalized else "<i4"
    h.update(np.asarray(obs["player_1"]).astype(dt).tobytes())
    h.update(np.asarray(obs["player_2"]).astype(dt).tobytes())


def run_session(args):
    """One reference env (with the reference's own wrappers), checked call by call against the C
    oracle driven through its batched entry point with n = 1 (pk_vec_step_ex, auto-reset on)."""
    cfg, action_mode, seed, env_index, episodes, max_calls = args
    env = rh.make_env(seed, **cfg)
    raw = env.raw
    n_actions = 13 if cfg.get("simplify_action") else 18
    normalized = bool(cfg.get("normalize_observation"))
    record = bool(cfg.get("record_episode_statistics"))

    orc = po.OracleVecEnv(1, seed=seed, autoreset=True, **cfg)

    def check(obs, rew, term, where):
        ref_obs = np.concatenate([obs["player_1"], obs["player_2"]])
        mine = orc.normalized_obs().reshape(70) if normalized else orc.obs.reshape(70)
        if ref_obs.dtype != (np.float64 if normalized else np.int64) or not np.array_equal(ref_obs, mine):
            raise AssertionError(f"obs mismatch {where}: {np.nonzero(ref_obs != mine)[0]}")
        ref_state = rh.unpacked_state(raw)
        if not np.array_equal(ref_state, orc.state[0, :52]):
            raise AssertionError(f"state mismatch {where}: words {np.nonzero(ref_state != orc.state[0, :52])[0]}")
        if rew is not None:
            r = np.array([rew["player_1"], rew["player_2"]], dtype=np.float64)
            if not np.array_equal(r, orc.reward[0]) or bool(term) != bool(orc.done[0]):
                raise AssertionError(f"reward/term mismatch {where}: {r} {orc.reward[0]} {term} {orc.done[0]}")

    h = hashlib.sha256()
    obs, _ = env.reset()
    orc.reset()
    check(obs, None, None, (seed, "reset0"))
    _hash_obs(h, obs, normalized)

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
        obs, rew, terms, _, infos = env.step({"player_1": a1, "player_2": a2})
        orc.step(np.array([[a1, a2]], dtype=np.int32))
        calls += 1
        frame += 1
        ep_frames += 1
        term = bool(terms["player_1"])
        check(obs, rew, term, (seed, len(results), ep_frames))
        _hash_obs(h, obs, normalized)
        h.update(np.array([rew["player_1"], rew["player_2"]], dtype="<f8").tobytes())
        h.update(bytes([int(term)]))
        if term:
            ep = {"frames": ep_frames, "scores": [int(raw.scores[0]), int(raw.scores[1])]}
            if record:  # RecordEpisodeStatistics, record_episode_statistics.py:34-39
                r = [infos[a]["episode"]["r"] for a in ("player_1", "player_2")]
                ln = [infos[a]["episode"]["l"] for a in ("player_1", "player_2")]
                assert ln[0] == ln[1] == ep_frames == int(orc.episode_length[0])
                if not np.array_equal(np.array(r, dtype=np.float64), orc.episode_return[0]):
                    raise AssertionError(f"episode return mismatch {seed}: {r} {orc.episode_return[0]}")
                h.update(np.array(r, dtype="<f8").tobytes())
                h.update(np.array([ln[0]], dtype="<i4").tobytes())
                ep["returns"] = [float(r[0]), float(r[1])]
            results.append(ep)
            ep_frames = 0
            if len(results) < episodes and calls < max_calls:
                obs, _ = env.reset()  # the product's auto-reset call
                orc.step(np.array([[a1, a2]], dtype=np.int32))  # a call on a terminated env = reset()
                calls += 1
                frame += 1
                check(obs, None, None, (seed, len(results), "reset"))
                if record:
                    assert orc.episode_return[0].tolist() == [0.0, 0.0] and orc.episode_length[0] == 0
                _hash_obs(h, obs, normalized)
    return {
        "seed": seed,
        "episodes": results,
        "calls": calls,
        "unfinished_frames": ep_frames,
        "sha256": h.hexdigest(),
        "final_state": [int(v) for v in orc.state[0, :52]],
    }


def run_orders(out_path):
    """Wrapper stacks in orders the fused kernel options do not cover (oracle/wrapper_orders.py): the reference's
    own wrapper classes, stacked as listed, played with seeded actions -> tests/golden/wrapper_orders.json."""
    from oracle import wrapper_orders as wo

    pikazoo_v0, W = rh.load_reference()
    out = {"generator": "oracle/make_golden.py --orders", "reference": "helpingstar/pika-zoo @ /root/reference "
           "(unmodified), numpy " + np.__version__, "env": wo.ENV_KW, "action_seed": wo.ACTION_SEED, "stacks": {}}
    for name, stack in wo.STACKS.items():
        sessions = []
        for seed in wo.SEEDS:
            raw = pikazoo_v0.env(**wo.ENV_KW)
            raw.np_random.bit_generator.state = np.random.PCG64(int(seed)).state  # protocol S0
            env = wo.build_stack(raw, W, stack)
            n_actions = env.action_space("player_1").n
            res = wo.run_session(env, n_actions, lambda f, a: synth_action(wo.ACTION_SEED, seed, f, a, n_actions))
            sessions.append(dict(seed=seed, **res))
        out["stacks"][name] = {"stack": [list(map(lambda v: list(v) if isinstance(v, tuple) else v, item))
                                         for item in stack], "sessions": sessions}
        print(name, [len(s["episodes"]) for s in sessions], [s["calls"] for s in sessions], flush=True)
    with open(out_path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1024:.0f} KiB")


RENDER_SESSIONS = [
    # (name, env kwargs, action mode, seed, cloud seed, frames)
    ("random_ws3", dict(winning_score=3, serve="random"), "synth", 31, 900, 700),
    ("ai_vs_ai_ws2", dict(winning_score=2, serve="winner", is_player1_computer=True, is_player2_computer=True), "noop",
     32, 901, 1500),
    ("ai_vs_random_ws12", dict(winning_score=12, serve="alternate", is_player2_computer=True), "synth", 33, 902, 2500),
]


def item_hash(items):
    h = hashlib.sha256()
    for name, flip, w, h_, x, y in items:
        h.update(name.encode())
        h.update(np.array([flip, w, h_, x, y], dtype="<i4").tobytes())
    return h.hexdigest()[:16]


def run_render(out_path):
    """Display lists of the reference's own draw() (pikazoo_env.py:250-384) recorded through the pygame stand-in of
    oracle/pygame_stub.py -> tests/golden/render.json. The clouds and the wave draw from their own generator
    (env.np_random is swapped around every render() call, and the ten Cloud objects are re-created from it with the
    reference's class), so that the game stream stays the S0 stream; zero-sized blits are dropped."""
    from oracle import pygame_stub  # noqa: F401  (installed by ref_harness)

    pikazoo_v0, _ = rh.load_reference()
    import pikazoo.env.cloud_and_wave as caw  # the reference's module

    out = {"generator": "oracle/make_golden.py --render", "reference": "helpingstar/pika-zoo @ /root/reference "
           "(unmodified draw(), pygame replaced by oracle/pygame_stub.py), numpy " + np.__version__,
           "item": "[sprite file, x-flipped, width, height, x, y]", "action_seed": ACTION_SEED, "sessions": []}
    for name, kw, mode, seed, cloud_seed, frames in RENDER_SESSIONS:
        env = pikazoo_v0.env(render_mode="rgb_array", **kw)
        game_rng = env.np_random
        game_rng.bit_generator.state = np.random.PCG64(int(seed)).state  # protocol S0
        cloud_rng = np.random.Generator(np.random.PCG64(int(cloud_seed)))
        env.cloud_array = [caw.Cloud(cloud_rng) for _ in range(env.NUM_OF_CLOUDS)]

        def render():
            env.np_random = cloud_rng
            frame = env.render()
            env.np_random = game_rng
            assert frame.shape == (304, 432, 3)
            log = [it for it in env.screen.log if it[2] > 0 and it[3] > 0]
            del env.screen.log[:]
            return log

        env.reset()
        first = render()
        n_static = next(i for i, it in enumerate(first) if it[0] == "cloud.png")  # draw_background's blits
        background = first[:n_static]
        hashes, samples, lengths = [item_hash(first[n_static:])], {0: first[n_static:]}, [len(first) - n_static]
        kinds = set()
        for f in range(frames):
            if mode == "noop":
                a1 = a2 = 0
            else:
                a1, a2 = synth_action(ACTION_SEED, seed, f, 0, 18), synth_action(ACTION_SEED, seed, f, 1, 18)
            _, _, term, _, _ = env.step({"player_1": a1, "player_2": a2})
            if term["player_1"]:
                env.reset()
            log = render()
            assert log[:n_static] == background
            dyn = log[n_static:]
            hashes.append(item_hash(dyn))
            lengths.append(len(dyn))
            new = {it[0] for it in dyn} - kinds
            if new or (f + 1) % 250 == 0:   # keep the full list whenever a sprite appears for the first time
                samples[f + 1] = dyn
            kinds |= new
        out["sessions"].append({"name": name, "config": kw, "action_mode": mode, "seed": seed, "cloud_seed": cloud_seed,
                                "frames": frames, "scores": [int(v) for v in env.scores], "hashes": hashes,
                                "lengths": lengths, "samples": {str(k): v for k, v in samples.items()},
                                "sprites_seen": sorted(kinds)})
        print(name, "frames", frames, "scores", env.scores, "sprites", len(kinds), "max items", max(lengths), flush=True)
        out["background"] = background
    with open(out_path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1024:.0f} KiB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--render", action="store_true", help="display lists of draw() -> tests/golden/render.json")
    ap.add_argument("--orders", action="store_true", help="odd wrapper orders -> tests/golden/wrapper_orders.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--wrappers", action="store_true", help="the wrapper-stack groups -> tests/golden/wrappers.json")
    ap.add_argument("--out", default=None)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    a = ap.parse_args()
    if not rh.reference_available():
        raise SystemExit("reference not present; golden fixtures can only be generated in the build container")
    po.build()
    if a.render:
        run_render(a.out or os.path.join(_ROOT, "tests", "golden", "render.json"))
        return
    if a.orders:
        run_orders(a.out or os.path.join(_ROOT, "tests", "golden", "wrapper_orders.json"))
        return
    groups = GROUPS_WRAPPERS if a.wrappers else (GROUPS_QUICK if a.quick else GROUPS_FULL)
    if a.out is None:
        a.out = os.path.join(_ROOT, "tests", "golden", "wrappers.json" if a.wrappers else "sessions.json")
    out = {
        "generator": "oracle/make_golden.py",
        "reference": "helpingstar/pika-zoo @ /root/reference (unmodified), numpy " + np.__version__,
        "seeding": "S0: env.np_random.bit_generator.state = np.random.PCG64(seed).state; reset()",
        "action_seed": ACTION_SEED,
        "hash": "sha256(reset obs | per step: obs1, obs2 int32 LE (float64 LE when normalize_observation), rewards "
                "2xfloat64 LE, terminated byte, [on termination with record_episode_statistics: episode returns "
                "2xfloat64 LE, episode length int32 LE] | reset obs after a terminated step)",
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
