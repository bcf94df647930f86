"""TEST / MEASUREMENT INFRASTRUCTURE — times the UNMODIFIED Python reference on this machine's cores.

BASELINE.md's CPU-baseline plan: one env process per core, each stepping the reference
(`/root/reference`, imported through oracle/ref_harness.py's stand-ins, seeded by protocol S0) on the
bench workload (uniform random Discrete(18) actions, winning_score 15, serve "winner", reset() on
termination) or on configs[0] (computer vs computer, NOOP actions) for a fixed wall time. The reference
tree does not exist on the GPU box, so this can only run in the build container; its output is committed
as profiles/r01_python_reference_cpu_container.json and quoted next to the C port's numbers.

    python -m oracle.time_reference --seconds 10
"""

from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import ref_harness as rh  # noqa: E402


def _worker(args):
    idx, seconds, ai = args
    kw = dict(winning_score=15, serve="winner", is_player1_computer=ai, is_player2_computer=ai)
    env = rh.make_env(1000 + idx, **kw)
    env.reset()
    rng = np.random.default_rng(idx)
    acts = rng.integers(0, 18, size=(4096, 2))
    # warm-up
    t_end = time.perf_counter() + 1.0
    k = 0
    while time.perf_counter() < t_end:
        a = acts[k & 4095]
        _, _, term, _, _ = env.step({"player_1": 0 if ai else int(a[0]), "player_2": 0 if ai else int(a[1])})
        if term["player_1"]:
            env.reset()
        k += 1
    steps = 0
    t0 = time.perf_counter()
    t_end = t0 + seconds
    while time.perf_counter() < t_end:
        for _ in range(256):
            a = acts[k & 4095]
            _, _, term, _, _ = env.step({"player_1": 0 if ai else int(a[0]), "player_2": 0 if ai else int(a[1])})
            if term["player_1"]:
                env.reset()
            k += 1
        steps += 256
    return steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--out", default=os.path.join(_ROOT, "profiles", "r01_python_reference_cpu_container.json"))
    a = ap.parse_args()
    if not rh.reference_available():
        raise SystemExit("reference tree not present (build container only)")
    cores = len(os.sched_getaffinity(0))
    out = {"what": "unmodified pure-Python reference (helpingstar/pika-zoo) in the BUILD CONTAINER, one env process "
                   "per core, protocol S0 seeding, gymnasium/pettingzoo/pygame replaced by oracle/ref_harness.py "
                   "stand-ins, render_mode=None", "cores": cores, "seconds": a.seconds, "numpy": np.__version__,
           "python": sys.version.split()[0]}
    with mp.Pool(cores) as pool:
        for name, ai in (("random_vs_random_ws15_winner (bench workload)", False),
                         ("computer_vs_computer_ws15_winner (configs[0] / configs[3])", True)):
            rates = pool.map(_worker, [(i, a.seconds, ai) for i in range(cores)])
            out[name] = {"env_steps_per_sec_total": sum(rates), "per_core_mean": sum(rates) / cores,
                         "per_core": [round(r) for r in rates]}
            print(name, f"{sum(rates):.0f} env-steps/s over {cores} cores ({sum(rates) / cores:.0f} per core)")
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
