"""e2e of the host-buffer path (pz_host_step, reference dtypes: int32 obs [n,2,35], float32 reward, uint8 done) per
wire format: native, compact (pz_host_set_wire) and the hybrid split (the last `native_chunks` chunks travel natively
while host threads rebuild the others), over chunk counts and thread counts.
    python profiles/time_host_wire.py [n] > gpurun_out/host_wire.json"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pikazoo_b200  # noqa: E402
from pikazoo_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
L = _lib.load()
torch.cuda.set_device(0)
h_act = [torch.randint(0, 18, (n, 2), dtype=torch.int32).pin_memory() for _ in range(2)]
h_obs = torch.empty((n, 2, 35), dtype=torch.int32).pin_memory()
h_rew = torch.empty((n, 2), dtype=torch.float32).pin_memory()
h_done = torch.empty((n,), dtype=torch.uint8).pin_memory()


def run(chunks, wire, threads, native_chunks, steps=30):
    cfg = pikazoo_b200.make_config(winning_score=15, serve="winner")
    ctx = ctypes.c_void_p()
    _lib.check(L.pz_host_create(ctypes.byref(ctx), n, ctypes.byref(cfg), 2026, 0, chunks), "create")
    if wire:
        _lib.check(L.pz_host_set_wire(ctx, 1, threads), "set_wire")
        if native_chunks:  # the hybrid split existed as an experiment only (see pz_host.cu); rows recorded with it are
            raise SystemExit("pz_host_set_wire_split was removed: rerun with native_chunks = 0")  # in r02_host_wire_sweep.json
    _lib.check(L.pz_host_reset(ctx, h_obs.data_ptr()), "reset")
    for k in range(3):
        _lib.check(L.pz_host_step(ctx, h_act[k % 2].data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr()), "step")
    t0 = time.perf_counter()
    for k in range(steps):
        _lib.check(L.pz_host_step(ctx, h_act[k % 2].data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(), h_done.data_ptr()), "step")
    dt = (time.perf_counter() - t0) / steps
    L.pz_host_destroy(ctx)
    return {"chunks": chunks, "wire": wire, "threads": threads, "native_chunks": native_chunks,
            "ms_per_step": round(dt * 1e3, 3), "M_env_steps_per_s": round(n / dt / 1e6, 1)}


cores = len(os.sched_getaffinity(0))
rows = [run(8, 0, 0, 0)]
for chunks, natives in ((8, (0,)), (16, (0,)), (32, (0,))):
    for nat in natives:
        for th in (cores, cores - 2):
            rows.append(run(chunks, 1, th, nat))
            print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
print(json.dumps({"n": n, "cores": cores, "rows": rows}))
