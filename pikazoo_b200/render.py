"""Batched `rgb_array` rendering of selected envs: the reference's `raw_env.render()`
(pikazoo/env/pikazoo_env.py:250-384) for a handful of envs of a batch, frames produced on the device.

What the reference's renderer needs beyond the simulation state is render-only state that the hot kernels do not
carry: the ball's `fine_rotation` / `rotation` (physics.py:373-388), `punch_effect_radius` /
`punch_effect_y` (physics.py:275,429-430,629-632; decremented by draw_ball itself, pikazoo_env.py:296-297) and the
cloud / wave animation (cloud_and_wave.py). All of it is a function of consecutive simulation states, so it is
tracked HERE, on the host, for the selected envs only (`RenderTracker.on_step(pre, post)` from the exported states
around every call) — the step kernels are untouched. The clouds and waves draw from their OWN numpy generator
(`cloud_seed`): in the reference they share the game's generator, so that merely calling render() changes the game;
here rendering never touches the game stream.

Pipeline per frame: exported state -> display list (which sprite where, in the reference's draw order; pinned
against the unmodified reference's own draw() by tests/test_render.py through a recording pygame stand-in) ->
`pz_render` (csrc/pz_render.cu), one thread per pixel walking the list back to front. Every sprite of the
reference has binary alpha, so compositing is exact; the two scaled sprites (clouds, punch effect) use the
nearest-neighbour mapping of pygame.transform.scale as restated in `scale_indices` (pygame is not installable here:
that mapping is unpinned).

The sprites are the reference's assets and are not part of this repository: pass `sprite_dir` (the `img/` directory
of an installed pikazoo), or set PIKAZOO_SPRITE_DIR.
"""

from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

WIDTH, HEIGHT = 432, 304  # GROUND_WIDTH, GROUND_HEIGHT (pikazoo_env.py:24)
NUM_CLOUDS, NUM_WAVES = 10, 432 // 16
MAX_ITEMS = 64

# pikazoo_env.py:437-466, the order of the `pikachu` tuple
PIKACHU_FILES = tuple(f"pikachu_{s}_{f}.png" for s, n in ((0, 5), (1, 5), (2, 5), (3, 2), (4, 1), (5, 5), (6, 5))
                      for f in range(n))
BALL_FILES = ("ball_0.png", "ball_1.png", "ball_2.png", "ball_3.png", "ball_4.png", "ball_hyper.png")
STATIC_FILES = ("sky_blue.png", "mountain.png", "ground_red.png", "ground_line.png", "ground_line_leftmost.png",
                "ground_line_rightmost.png", "ground_yellow.png", "net_pillar_top.png", "net_pillar.png")
DYNAMIC_FILES = PIKACHU_FILES[:18] + BALL_FILES + ("ball_trail.png", "ball_punch.png", "shadow.png", "cloud.png",
                                                   "wave.png") + tuple(f"number_{d}.png" for d in range(10))

Item = Tuple[str, int, int, int, int, int]  # (sprite file, x-flipped, width, height, x, y) — one blit


# ---- sprites ---------------------------------------------------------------------------------------------------
def decode_png(path: str) -> np.ndarray:
    """uint8 [h, w, 4] of an 8-bit RGBA, non-interlaced PNG (all 70 sprites of the reference are)."""
    with open(path, "rb") as f:
        d = f.read()
    if d[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG")
    p, idat, w, h = 8, b"", 0, 0
    while p < len(d):
        (length,) = struct.unpack(">I", d[p:p + 4])
        kind, body = d[p + 4:p + 8], d[p + 8:p + 8 + length]
        p += 12 + length
        if kind == b"IHDR":
            w, h, depth, colour, _, _, interlace = struct.unpack(">IIBBBBB", body)
            if (depth, colour, interlace) != (8, 6, 0):
                raise ValueError(f"{path}: only 8-bit RGBA non-interlaced PNGs are supported")
        elif kind == b"IDAT":
            idat += body
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    out = np.zeros((h, 4 * w), dtype=np.uint8)
    prev = np.zeros(4 * w, dtype=np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:  # filters with a left neighbour: sequential over the pixels, vectorised over the four channels
            cur = np.zeros(4 * w, dtype=np.int32)
            for i in range(0, 4 * w, 4):
                a = cur[i - 4:i] if i else np.zeros(4, dtype=np.int32)
                b = prev[i:i + 4]
                c = prev[i - 4:i] if i else np.zeros(4, dtype=np.int32)
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                elif ft == 4:
                    pa, pb, pc = np.abs(b - c), np.abs(a - c), np.abs(a + b - 2 * c)
                    pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
                else:
                    raise ValueError(f"{path}: bad filter {ft}")
                cur[i:i + 4] = (line[i:i + 4] + pred) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 4)


def scale_indices(src: int, dst: int) -> np.ndarray:
    """Source index of every destination pixel of pygame.transform.scale along one axis (its `stretch` loop: a
    Bresenham walk over the source). Unpinned restatement (pygame is not installable here)."""
    idx, out, err = 0, np.zeros(dst, dtype=np.int64), 2 * src - 2 * dst
    for i in range(dst):
        out[i] = min(idx, src - 1)
        while err >= 0:
            idx += 1
            err -= 2 * dst
        err += 2 * src
    return out


def scale_image(img: np.ndarray, w: int, h: int) -> np.ndarray:
    if w <= 0 or h <= 0:
        return np.zeros((0, 0, 4), dtype=np.uint8)
    return img[scale_indices(img.shape[0], h)][:, scale_indices(img.shape[1], w)]


class SpriteSet:
    """The sprites the renderer draws, decoded from `sprite_dir`, with every variant the display lists can name
    (x-flipped players, the six cloud sizes, the punch-effect sizes), the static background precomposed."""

    def __init__(self, sprite_dir: Optional[str] = None, images: Optional[Dict[str, np.ndarray]] = None):
        if images is None:
            sprite_dir = sprite_dir or os.environ.get("PIKAZOO_SPRITE_DIR") or _find_installed_sprites()
            if not sprite_dir or not os.path.isdir(sprite_dir):
                raise FileNotFoundError(
                    "rendering needs the reference's sprite directory (pikazoo/env/img of helpingstar/pika-zoo; the "
                    "PNGs are not part of this repository): pass sprite_dir= or set PIKAZOO_SPRITE_DIR")
            images = {f: decode_png(os.path.join(sprite_dir, f)) for f in STATIC_FILES + DYNAMIC_FILES}
        for f, im in images.items():
            a = np.unique(im[..., 3])
            if not set(a.tolist()) <= {0, 255}:
                raise ValueError(f"{f}: alpha must be binary (the compositing here is a select, not a blend)")
        self.images = images
        self.size = {f: (im.shape[1], im.shape[0]) for f, im in images.items()}
        self.background_items = static_items(self.size)
        self.background = composite(self.background_items, self, np.zeros((HEIGHT, WIDTH, 3), dtype=np.uint8))
        self._variants: Dict[Tuple[str, int, int, int], int] = {}
        self._pixels: List[np.ndarray] = []
        self._table: List[Tuple[int, int, int]] = []  # (offset in RGBA texels, w, h)
        self._offset = 0
        for f in DYNAMIC_FILES:  # every variant a display list can name, so that the atlas is built once
            w, h = self.size[f]
            self.variant(f, 0, w, h)
            if f.startswith("pikachu"):
                self.variant(f, 1, w, h)
        for sd in range(6):
            self.variant("cloud.png", 0, 48 + 2 * sd, 24 + 2 * sd)
        for r in range(2, 20, 2):
            self.variant("ball_punch.png", 0, 2 * r, 2 * r)

    def pixels(self, name: str, flip: int, w: int, h: int) -> np.ndarray:
        im = self.images[name]
        if (w, h) != self.size[name]:
            im = scale_image(im, w, h)
        return im[:, ::-1] if flip else im

    def variant(self, name: str, flip: int, w: int, h: int) -> int:
        key = (name, int(flip), int(w), int(h))
        if key not in self._variants:
            px = np.ascontiguousarray(self.pixels(*key))
            self._variants[key] = len(self._table)
            self._table.append((self._offset, px.shape[1], px.shape[0]))
            self._pixels.append(px.reshape(-1, 4))
            self._offset += px.shape[0] * px.shape[1]
        return self._variants[key]

    def atlas(self) -> Tuple[np.ndarray, np.ndarray]:
        """(uint8 [texels, 4], int32 [variants, 4] = offset, w, h, 0) for the device."""
        table = np.zeros((len(self._table), 4), dtype=np.int32)
        table[:, :3] = np.array(self._table, dtype=np.int32)
        return np.concatenate(self._pixels, axis=0), table


def _find_installed_sprites() -> Optional[str]:
    try:
        import importlib.util

        spec = importlib.util.find_spec("pikazoo")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "env", "img")
            return cand if os.path.isdir(cand) else None
    except Exception:  # noqa: BLE001
        pass
    return None


# ---- display lists (pikazoo_env.py:250-384) ----------------------------------------------------------------------
def static_items(size: Dict[str, Tuple[int, int]]) -> List[Item]:
    """draw_background, pikazoo_env.py:308-339: the same 441 blits every frame."""
    def it(name, x, y):
        return (name, 0, size[name][0], size[name][1], x, y)

    out = [it("sky_blue.png", 16 * i, 16 * j) for j in range(12) for i in range(WIDTH // 16)]
    out.append(it("mountain.png", 0, 188))
    out += [it("ground_red.png", 16 * i, 248) for i in range(WIDTH // 16)]
    out += [it("ground_line.png", 16 * i, 264) for i in range(1, WIDTH // 16 - 1)]
    out += [it("ground_line_leftmost.png", 0, 264), it("ground_line_rightmost.png", WIDTH - 16, 264)]
    out += [it("ground_yellow.png", 16 * i, 280 + 16 * j) for j in range(2) for i in range(WIDTH // 16)]
    out.append(it("net_pillar_top.png", 213, 176))
    out += [it("net_pillar.png", 213, 184 + 8 * j) for j in range(12)]
    return out


def player_sprite_index(state: int, frame: int) -> int:
    """get_frame_number_for_player_animated_sprite, pikazoo_env.py:46-69"""
    if state < 4:
        return 5 * state + frame
    if state == 4:
        return 17 + frame
    return 18 + 5 * (state - 5) + frame


class CloudsAndWave:
    """Cloud / Wave / cloud_and_wave_engine (cloud_and_wave.py:11-78) of one env, on its own generator."""

    def __init__(self, rng: np.random.Generator):
        self.rng = rng
        self.clouds = []
        for _ in range(NUM_CLOUDS):  # Cloud.__init__, :14-18 (draw order matters)
            x = -68 + int(rng.integers(0, 432 + 68))
            y = int(rng.integers(0, 152))
            v = 1 + int(rng.integers(0, 2))
            turn = int(rng.integers(0, 11))
            self.clouds.append([x, y, v, turn])
        self.vertical_coord, self.vertical_velocity = 0, 2
        self.y_coords = [314] * NUM_WAVES

    def advance(self) -> None:
        rng = self.rng
        for c in self.clouds:  # :58-64
            c[0] += c[2]
            if c[0] > 432:
                c[0] = -68
                c[1] = int(rng.integers(0, 152))
                c[2] = 1 + int(rng.integers(0, 2))
            c[3] = (c[3] + 1) % 11
        self.vertical_coord += self.vertical_velocity  # :66-72
        if self.vertical_coord > 32:
            self.vertical_coord = 32
            self.vertical_velocity = -1
        elif self.vertical_coord < 0 and self.vertical_velocity < 0:
            self.vertical_velocity = 2
            self.vertical_coord = -int(rng.integers(0, 40))
        for i in range(NUM_WAVES):  # :74-75
            self.y_coords[i] = 314 - self.vertical_coord + int(rng.integers(0, 3))

    def items(self) -> List[Item]:
        """draw_clouds_and_wave after the engine ran, pikazoo_env.py:353-366"""
        out = []
        for x, y, _, turn in self.clouds:
            sd = 5 - abs(turn - 5)  # Cloud.size_diff, :21-22
            out.append(("cloud.png", 0, 48 + 2 * sd, 24 + 2 * sd, x - sd, y - sd))
        out += [("wave.png", 0, 16, 32, 16 * i, self.y_coords[i]) for i in range(NUM_WAVES)]
        return out


class RenderTracker:
    """The render-only ball state of one env, maintained from the unpacked states (oracle/pika_oracle.h pk_env
    word order, `PikaVecEnv.export_state`) before and after every call of the env."""

    def __init__(self):
        self.fine_rotation = 0      # physics.py:238
        self.rotation = 0           # :237
        self.punch_radius = 0       # :275
        self.punch_y = 0            # :242

    def on_reset(self) -> None:
        """raw_env.reset -> Ball.initialize_for_new_round (physics.py:275); the rotation carries over"""
        self.punch_radius = 0

    def on_step(self, pre: Sequence[int], post: Sequence[int]) -> None:
        """One executed frame: pre / post = the env's 53 words before / after it."""
        new_round = bool(pre[39])                  # round_ended: this call re-initialised the round first
        if new_round:
            self.punch_radius = 0                  # physics.py:275
        xv = 0 if new_round else int(pre[28])      # ball x velocity entering the world-collision step
        f = self.fine_rotation + (xv // 2)         # :373 (floor division)
        if f < 0:
            f += 50
        elif f > 50:
            f -= 50
        self.fine_rotation, self.rotation = f, f // 10   # :387-388 (50 -> rotation 5: the "hyper ball glitch")
        if post[39]:                               # the ball touched the ground in this frame, :428-430
            self.punch_radius, self.punch_y = 20, 252 + 20
        for k in (0, 13):                          # a NEW collision with a power-hitting player, :629-632
            was = 0 if new_round else int(pre[k + 9])
            if post[k + 9] and not was and post[k + 3] == 2:
                self.punch_radius, self.punch_y = 20, int(post[27])


def dynamic_items(state: Sequence[int], tr: RenderTracker, cw: CloudsAndWave, size: Dict[str, Tuple[int, int]]) -> List[Item]:
    """Everything draw() blits after the background, in its order (pikazoo_env.py:250-306,341-366). Mutates the
    tracker exactly as draw_ball mutates the ball (the punch effect shrinks by 2 per rendered frame), and advances the
    clouds and the wave by one engine step, as draw_clouds_and_wave does. Zero-sized blits are dropped."""
    cw.advance()
    out = cw.items()
    flips = []
    for k, flip_dir in ((0, -1), (13, 1)):
        st, frame, dive = int(state[k + 3]), int(state[k + 4]), int(state[k + 7])
        diving = st in (3, 4) and dive == flip_dir
        flips.append(diving if k == 0 else not diving)          # :262-263
        name = PIKACHU_FILES[player_sprite_index(st, frame)]
        w, h = size[name]
        out.append((name, int(flips[-1]), w, h, int(state[k]) - w // 2, int(state[k + 1]) - h // 2))
    sw, sh = size["shadow.png"]
    out += [("shadow.png", 0, sw, sh, int(state[k]) - sw // 2, 273 - sh // 2) for k in (0, 13)]
    bx, by, px, py, ppx, ppy, power, punch_x = (int(state[26]), int(state[27]), int(state[30]), int(state[31]),
                                                int(state[32]), int(state[33]), int(state[34]), int(state[36]))
    name = BALL_FILES[tr.rotation]
    w, h = size[name]
    out.append((name, 0, w, h, bx - w // 2, by - h // 2))
    out.append(("shadow.png", 0, sw, sh, bx - sw // 2, 273 - sh // 2))
    if power:
        for name, cx, cy in (("ball_hyper.png", px, py), ("ball_trail.png", ppx, ppy)):
            w, h = size[name]
            out.append((name, 0, w, h, cx - w // 2, cy - h // 2))
    if tr.punch_radius > 0:                                       # :296-306
        tr.punch_radius -= 2
        r = tr.punch_radius
        if r > 0:
            out.append(("ball_punch.png", 0, 2 * r, 2 * r, punch_x - r, tr.punch_y - r))
    s1, s2 = int(state[37]), int(state[38])                        # :341-351
    nw = 32
    if s1 >= 10:
        out.append(("number_1.png", 0, *size["number_1.png"], 14, 10))
    out.append((f"number_{s1 % 10}.png", 0, *size[f"number_{s1 % 10}.png"], 14 + nw, 10))
    if s2 >= 10:
        out.append(("number_1.png", 0, *size["number_1.png"], WIDTH - 32 - 32 - 14, 10))
    out.append((f"number_{s2 % 10}.png", 0, *size[f"number_{s2 % 10}.png"], WIDTH - 32 - 32 - 14 + nw, 10))
    return out


def composite(items: Sequence[Item], sprites: SpriteSet, canvas: np.ndarray) -> np.ndarray:
    """numpy rasteriser of a display list onto an RGB canvas [HEIGHT, WIDTH, 3] (front = later items): builds the
    static background, and is what tests compare the CUDA rasteriser with."""
    for name, flip, w, h, x, y in items:
        px = sprites.pixels(name, flip, w, h)
        x0, y0, x1, y1 = max(x, 0), max(y, 0), min(x + w, WIDTH), min(y + h, HEIGHT)
        if x0 >= x1 or y0 >= y1:
            continue
        src = px[y0 - y:y1 - y, x0 - x:x1 - x]
        dst = canvas[y0:y1, x0:x1]
        mask = src[..., 3] != 0
        dst[mask] = src[..., :3][mask]
    return canvas


# ---- the batch renderer ------------------------------------------------------------------------------------------
class BatchRenderer:
    """rgb_array frames of `indices` of a PikaVecEnv. Attach it BEFORE stepping (`env.attach_renderer(...)`): every
    `env.reset()` / `env.step()` then reports the selected envs' states to it; `render()` returns a uint8 CUDA tensor
    [len(indices), 304, 432, 3] for the current states (call it once per step for the reference's animation speed:
    clouds, waves and the punch effect advance per rendered frame, as in the reference)."""

    def __init__(self, env, indices: Sequence[int], sprite_dir: Optional[str] = None, cloud_seed: int = 0,
                 sprites: Optional[SpriteSet] = None):
        import torch

        from . import _lib

        self.env, self.lib = env, _lib.load()
        self.indices = [int(i) for i in indices]
        self.sprites = sprites or SpriteSet(sprite_dir)
        self.trackers = [RenderTracker() for _ in self.indices]
        self.clouds = [CloudsAndWave(np.random.Generator(np.random.PCG64(int(cloud_seed) + env.first_env + i)))
                       for i in self.indices]
        dev = env.device
        atlas, table = self.sprites.atlas()
        self._atlas = torch.from_numpy(atlas).to(dev)
        self._table = torch.from_numpy(table).to(dev)
        self._background = torch.from_numpy(self.sprites.background).to(dev)
        self._sel = torch.tensor(self.indices, dtype=torch.int64, device=dev)
        self._items = torch.zeros((len(self.indices), MAX_ITEMS, 4), dtype=torch.int32).pin_memory()
        self._items_dev = torch.zeros((len(self.indices), MAX_ITEMS, 4), dtype=torch.int32, device=dev)
        self.frames = torch.zeros((len(self.indices), HEIGHT, WIDTH, 3), dtype=torch.uint8, device=dev)
        self._state = self._export()
        self.last_items: List[List[Item]] = []

    def _export(self) -> np.ndarray:
        return self.env.export_state()[self._sel].cpu().numpy()

    # hooks called by PikaVecEnv
    def before_call(self) -> None:
        self._pre = self._export()

    def after_reset(self) -> None:
        for t in self.trackers:
            t.on_reset()
        self._state = self._export()

    def after_step(self) -> None:
        post = self._export()
        for t, pre, cur in zip(self.trackers, self._pre, post):
            over = pre[40] != 0 or (self.env.max_episode_frames > 0 and pre[52] >= self.env.max_episode_frames)
            if not over:
                t.on_step(pre, cur)
            elif cur[40] == 0 and cur[52] == 0:  # the call was the auto-reset of a finished episode
                t.on_reset()
        self._state = post

    def display_lists(self) -> List[List[Item]]:
        self.last_items = [dynamic_items(s, t, c, self.sprites.size)
                           for s, t, c in zip(self._state, self.trackers, self.clouds)]
        return self.last_items

    def render(self):
        import torch

        from . import _lib

        items = self._items.numpy()
        items[:] = -1
        for b, lst in enumerate(self.display_lists()):
            if len(lst) > MAX_ITEMS:
                raise RuntimeError("display list longer than MAX_ITEMS")
            for m, (name, flip, w, h, x, y) in enumerate(lst):
                items[b, m] = (self.sprites.variant(name, flip, w, h), x, y, 0)
        if self.sprites.atlas()[1].shape[0] != self._table.shape[0]:  # a variant the atlas did not have yet
            atlas, table = self.sprites.atlas()
            self._atlas, self._table = torch.from_numpy(atlas).to(self.env.device), torch.from_numpy(table).to(self.env.device)
        with torch.cuda.device(self.env.device):
            stream = torch.cuda.current_stream()
            self._items_dev.copy_(self._items, non_blocking=True)
            _lib.check(self.lib.pz_render(self._atlas.data_ptr(), self._table.data_ptr(), self._table.shape[0],
                                          self._background.data_ptr(), self._items_dev.data_ptr(), len(self.indices),
                                          MAX_ITEMS, self.frames.data_ptr(), stream.cuda_stream), "pz_render")
            stream.synchronize()  # the pinned item buffer is rewritten by the next call
        return self.frames
