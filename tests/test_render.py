"""Rendering. CPU part: the display lists — which sprite where, flipped or scaled how, in
which order — that pikazoo_b200.render derives from consecutive simulation states must equal, frame for frame, what
the UNMODIFIED reference's draw() (pikazoo_env.py:250-384) blits: tests/golden/render.json was recorded from it through
a recording pygame stand-in (oracle/pygame_stub.py, oracle/make_golden.py --render; 4,700 frames, three configs,
every sprite the renderer can draw incl. two-digit scores, the hyper-ball glitch, power-hit trails and punch effects).
The states come from the C oracle here, so this runs without a GPU. pygame itself is not installable, so the
RASTERISATION of a display list is unpinned against pygame; all sprites have binary alpha, which makes compositing a
select, and tests/test_gpu_render.py checks the CUDA rasteriser against a numpy one."""

import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle.synth import synth_action
from pikazoo_b200 import render as R
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def golden_render():
    with open(os.path.join(ROOT, "tests", "golden", "render.json")) as f:
        return json.load(f)


def item_hash(items):
    h = hashlib.sha256()
    for name, flip, w, h_, x, y in items:
        h.update(name.encode())
        h.update(np.array([flip, w, h_, x, y], dtype="<i4").tobytes())
    return h.hexdigest()[:16]


def sizes_from_golden(g):
    """sprite sizes as the reference loaded them (every sprite appears unscaled somewhere in the lists)"""
    size = {}
    for it in g["background"]:
        size[it[0]] = (it[2], it[3])
    for s in g["sessions"]:
        for lst in s["samples"].values():
            for name, flip, w, h, x, y in lst:
                if name not in ("cloud.png", "ball_punch.png"):
                    size[name] = (w, h)
    size.setdefault("cloud.png", (48, 24))
    size.setdefault("ball_punch.png", (40, 40))
    for k in range(5):  # ball rotations and digits not seen in a sample have the size of their siblings
        size.setdefault(f"ball_{k}.png", size["ball_0.png"])
    for d in range(10):
        size.setdefault(f"number_{d}.png", size["number_0.png"])
    for f in R.PIKACHU_FILES:
        size.setdefault(f, size["pikachu_0_0.png"])
    return size


def test_static_background_list(golden_render):
    size = sizes_from_golden(golden_render)
    assert [list(it) for it in R.static_items(size)] == golden_render["background"]


@pytest.mark.parametrize("k", [0, 1, 2])
def test_display_lists_match_the_reference_draw(golden_render, k):
    g = golden_render["sessions"][k]
    size = sizes_from_golden(golden_render)
    cfg = dict(g["config"])
    orc = po.OracleVecEnv(1, seed=g["seed"], autoreset=True, **cfg)
    tracker = R.RenderTracker()
    clouds = R.CloudsAndWave(np.random.Generator(np.random.PCG64(g["cloud_seed"])))
    orc.reset()
    tracker.on_reset()
    got = R.dynamic_items(orc.state[0], tracker, clouds, size)
    assert item_hash(got) == g["hashes"][0], (got, g["samples"]["0"])
    seen = set()
    for f in range(g["frames"]):
        a = (0, 0) if g["action_mode"] == "noop" else tuple(
            synth_action(golden_render["action_seed"], g["seed"], f, ag, 18) for ag in (0, 1))
        pre = orc.state[0].copy()
        orc.step(np.array([a], dtype=np.int32))
        tracker.on_step(pre, orc.state[0])
        if orc.done[0]:  # the reference session calls reset() right away
            orc.step(np.array([a], dtype=np.int32))
            tracker.on_reset()
        got = R.dynamic_items(orc.state[0], tracker, clouds, size)
        if item_hash(got) != g["hashes"][f + 1]:
            want = g["samples"].get(str(f + 1))
            raise AssertionError(f"frame {f + 1}: {[x for x in got if x[0] not in ('cloud.png', 'wave.png')]} vs {want}")
        assert len(got) == g["lengths"][f + 1] <= R.MAX_ITEMS
        seen |= {it[0] for it in got}
    assert sorted(seen) == g["sprites_seen"]
    if k == 2:
        assert "ball_hyper.png" in seen and "ball_trail.png" in seen and "ball_punch.png" in seen


def test_tracker_hyper_ball_glitch():
    """fine_rotation reaching exactly 50 selects the sixth ball sprite (physics.py:374-381)"""
    t = R.RenderTracker()
    pre = np.zeros(53, dtype=np.int32)
    post = np.zeros(53, dtype=np.int32)
    pre[28] = 20                      # x velocity 20 -> +10 per frame
    for _ in range(5):
        t.on_step(pre, post)
    assert (t.fine_rotation, t.rotation) == (50, 5)
    t.on_step(pre, post)
    assert (t.fine_rotation, t.rotation) == (10, 1)
    pre[28] = -3                      # floor division: -3 // 2 == -2
    t.on_step(pre, post)
    assert t.fine_rotation == 8


def test_scale_indices_and_png_decoder(tmp_path):
    import struct
    import zlib

    assert R.scale_indices(4, 4).tolist() == [0, 1, 2, 3]
    assert R.scale_indices(4, 8).tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert R.scale_indices(8, 4).tolist() == [0, 2, 4, 6]
    idx = R.scale_indices(40, 36)
    assert idx[0] == 0 and idx[-1] <= 39 and np.all(np.diff(idx) >= 1)
    # a PNG written here with every filter type round-trips through the decoder
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(7, 5, 4), dtype=np.uint8)
    rows, prev = [], np.zeros(20, dtype=np.int32)
    for y in range(7):
        cur = img[y].reshape(-1).astype(np.int32)
        ft = y % 5
        a = np.concatenate([np.zeros(4, dtype=np.int32), cur[:-4]])
        c = np.concatenate([np.zeros(4, dtype=np.int32), prev[:-4]])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = a
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (a + prev) >> 1
        else:
            pa, pb, pc = np.abs(prev - c), np.abs(a - c), np.abs(a + prev - 2 * c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        rows.append(bytes([ft]) + ((cur - pred) & 255).astype(np.uint8).tobytes())
        prev = cur

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body))

    data = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 5, 7, 8, 6, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(b"".join(rows))) + chunk(b"IEND", b""))
    path = tmp_path / "t.png"
    path.write_bytes(data)
    assert np.array_equal(R.decode_png(str(path)), img)
