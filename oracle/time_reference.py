"""TEST / MEASUREMENT INFRASTRUCTURE — times the UNMODIFIED Python reference on this machine's cores.

BASELINE.md's CPU-baseline plan and BASELINE.json's north_star: one env process per core, each stepping the
reference (`/root/reference` in the build container, the staged copy `oracle/_ref/` on the GPU box — see
oracle/stage_ref.py —, imported through oracle/ref_harness.py's stand-ins for gymnasium / pettingzoo / pygame,
seeded by protocol S0) on the bench workload (uniform random Discrete(18) actions, winning_score 15, serve
"winner", reset() on termination) or on configs[0] (computer vs computer, NOOP actions).

The worker processes are persistent: every sample is a fixed wall-clock window in which all of them step their
env, so a caller can take K samples back to back (bench.py --impl reference). bench.py runs this module in a
child interpreter (no CUDA context is ever forked).

    python -m oracle.time_reference --seconds 10                        # both workloads -> profiles/*.json
    python -m oracle.time_reference --json --workload bench --samples 20 --seconds 1 --warmup-seconds 2
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

WORKLOADS = {
    "bench": dict(ai=False, what="random_vs_random_ws15_winner (bench workload, configs[1] semantics)"),
    "ai": dict(ai=True, what="computer_vs_computer_ws15_winner (configs[0] / configs[3])"),
}


def _worker(conn, idx: int, ai: bool):
    kw = dict(winning_score=15, serve="winner", is_player1_computer=ai, is_player2_computer=ai)
    env = rh.make_env(1000 + idx, **kw)
    env.reset()
    acts = np.random.default_rng(idx).integers(0, 18, size=(4096, 2)).tolist()
    k = 0
    conn.send("ready")
    while True:
        seconds = conn.recv()
        if seconds is None:
            return
        steps = 0
        t0 = time.perf_counter()
        t_end = t0 + seconds
        while time.perf_counter() < t_end:
            for _ in range(256):
                a = acts[k & 4095]
                _, _, term, _, _ = env.step({"player_1": 0 if ai else a[0], "player_2": 0 if ai else a[1]})
                if term["player_1"]:
                    env.reset()
                k += 1
            steps += 256
        conn.send((steps, time.perf_counter() - t0))


class ReferencePool:
    """One process per core, each owning one reference env."""

    def __init__(self, workload: str = "bench", cores: int | None = None):
        if not rh.reference_available():
            raise RuntimeError(f"reference sources not present at {rh.REFERENCE_ROOT} (run oracle/stage_ref.py)")
        self.cores = cores or len(os.sched_getaffinity(0))
        self.workload = workload
        ctx = mp.get_context("fork")
        self.procs, self.conns = [], []
        for i in range(self.cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, i, WORKLOADS[workload]["ai"]), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            assert c.recv() == "ready"

    def sample(self, seconds: float):
        """All workers step for `seconds` of wall clock: (total env-steps/s, per-core rates)."""
        for c in self.conns:
            c.send(seconds)
        rates = []
        for c in self.conns:
            steps, dt = c.recv()
            rates.append(steps / dt)
        return sum(rates), rates

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=5)


def describe() -> dict:
    staged = os.path.join(rh.REFERENCE_ROOT, "STAGED.json")
    d = {"reference_root": rh.REFERENCE_ROOT, "numpy": np.__version__, "python": sys.version.split()[0]}
    if os.path.exists(staged):
        with open(staged) as f:
            files = json.load(f)["files"]
        d["staged_files"] = len(files)
        d["physics_py_sha256_16"] = files.get("pikazoo/env/physics.py", "")[:16]
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--samples", type=int, default=1)
    ap.add_argument("--warmup-seconds", type=float, default=1.0)
    ap.add_argument("--workload", choices=list(WORKLOADS) + ["both"], default="both")
    ap.add_argument("--json", action="store_true", help="print one JSON line and write no file")
    ap.add_argument("--out", default=os.path.join(_ROOT, "profiles", "r02_python_reference_cpu_container.json"))
    a = ap.parse_args()
    out = {"what": "unmodified pure-Python reference (helpingstar/pika-zoo), one env process per core, protocol S0 "
                   "seeding, gymnasium/pettingzoo/pygame replaced by oracle/ref_harness.py stand-ins, "
                   "render_mode=None", "seconds_per_sample": a.seconds, "samples": a.samples, **describe()}
    names = list(WORKLOADS) if a.workload == "both" else [a.workload]
    for name in names:
        pool = ReferencePool(name)
        out["cores"] = pool.cores
        if a.warmup_seconds > 0:
            pool.sample(a.warmup_seconds)
        totals, per_core = [], None
        for _ in range(max(1, a.samples)):
            total, per_core = pool.sample(a.seconds)
            totals.append(total)
        pool.close()
        out[name] = {"what": WORKLOADS[name]["what"], "env_steps_per_sec_total": sum(totals) / len(totals),
                     "samples_env_steps_per_sec": totals, "per_core_mean": sum(totals) / len(totals) / out["cores"],
                     "per_core_last_sample": [round(r) for r in per_core]}
        if not a.json:
            print(name, f"{out[name]['env_steps_per_sec_total']:.0f} env-steps/s over {out['cores']} cores "
                        f"({out[name]['per_core_mean']:.0f} per core)")
    if a.json:
        print(json.dumps(out), flush=True)
    else:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
