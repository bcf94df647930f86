"""Host-memory side of the compact wire format: how fast do T host threads rebuild the reference's arrays
(pz_wire_expand: 70 B in, 289 B out per env) on this box, next to a plain memset of the same output and to the
link's ~55 GB/s? Decides how many threads pz_host_set_wire wants. No GPU needed.
    python profiles/wire_expand_probe.py [n]"""
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pikazoo_b200 import _lib  # noqa: E402

L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rows = np.random.randint(-300, 500, size=(n, 35), dtype=np.int16)
st = np.random.randint(0, 3, size=(n,), dtype=np.uint8)
obs = np.zeros((n, 2, 35), np.int32)
rew = np.zeros((n, 2), np.float32)
done = np.zeros(n, np.uint8)
P = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
out = {"n": n, "cores": os.cpu_count(), "expand_ms": {}, "memset_ms": {}}
for T in (1, 2, 4, 8, 12, 16, 24, 32):
    if T > 2 * (os.cpu_count() or 1):
        break

    def work(t, what):
        lo, hi = n * t // T // 8 * 8, (n * (t + 1) // T // 8 * 8 if t + 1 < T else n)
        if what == "expand":
            L.pz_wire_expand(P(rows[lo:hi]), P(st[lo:hi]), hi - lo, 0, P(obs[lo:hi]), 0, P(rew[lo:hi]), P(done[lo:hi]))
        else:
            ctypes.memset(obs[lo:hi].ctypes.data, 1, (hi - lo) * 280)

    for what in ("expand", "memset"):
        best = 1e9
        for rep in range(5):
            ths = [threading.Thread(target=work, args=(t, what)) for t in range(T)]
            t0 = time.perf_counter()
            [x.start() for x in ths]
            [x.join() for x in ths]
            best = min(best, time.perf_counter() - t0)
        out[what + "_ms"][T] = round(best * 1e3, 3)
out["expand_out_gbs"] = {T: round(n * 289 / (ms * 1e-3) / 1e9, 1) for T, ms in out["expand_ms"].items()}
print(json.dumps(out))
