"""TEST INFRASTRUCTURE — a recording stand-in for the handful of pygame calls the reference's renderer makes
(pikazoo/env/pikazoo_env.py:31-43,250-384: image.load, Surface, Surface.blit, transform.flip / scale,
surfarray.pixels3d, init / quit). pygame is not installed here and cannot be (no network), so the reference's pixels
cannot be produced; what CAN be pinned is everything the reference's own draw() decides — which sprite goes where, in
which order, flipped or scaled to which size: every blit onto the screen surface is logged as
(sprite file name, x-flipped, width, height, x, y). oracle/make_golden.py --render records those display lists from
the unmodified reference; the product's renderer must reproduce them exactly (tests/test_render.py). Rasterising a
display list (all sprites have binary alpha, checked) is the product's own work and is tested separately."""

from __future__ import annotations

import os
import struct
import types

import numpy as np

SRCALPHA = 0x00010000


def _png_size(path):
    with open(path, "rb") as f:
        head = f.read(24)
    return struct.unpack(">II", head[16:24])


class Surface:
    def __init__(self, size, flags=0):
        self.size = (int(size[0]), int(size[1]))
        self.tag = None      # file name of the sprite this surface shows
        self.flipped = False
        self.log = []        # blits onto this surface

    def get_size(self):
        return self.size

    def get_width(self):
        return self.size[0]

    def get_height(self):
        return self.size[1]

    def blit(self, src, pos):
        if self.tag is None and not self.log and src.size == self.size and tuple(pos) == (0, 0):
            self.tag = src.tag  # get_image(): a fresh SRCALPHA surface receiving the loaded image IS that sprite
            self.flipped = src.flipped
            return
        self.log.append((src.tag, int(src.flipped), src.size[0], src.size[1], int(pos[0]), int(pos[1])))


def install(modules):
    """Put the stub into `modules` (sys.modules) as `pygame`."""
    pg = types.ModuleType("pygame")
    pg.SRCALPHA = SRCALPHA
    pg.Surface = Surface
    pg.init = lambda: None
    pg.quit = lambda: None

    image = types.ModuleType("pygame.image")

    def load(path):
        s = Surface(_png_size(path))
        s.tag = os.path.basename(path)
        return s

    image.load = load
    transform = types.ModuleType("pygame.transform")

    def flip(surface, xbool, ybool):
        assert xbool and not ybool
        s = Surface(surface.size)
        s.tag, s.flipped = surface.tag, not surface.flipped
        return s

    def scale(surface, size):
        s = Surface(size)
        s.tag, s.flipped = surface.tag, surface.flipped
        return s

    transform.flip, transform.scale = flip, scale
    surfarray = types.ModuleType("pygame.surfarray")
    surfarray.pixels3d = lambda surface: np.zeros((surface.size[0], surface.size[1], 3), dtype=np.uint8)
    pg.image, pg.transform, pg.surfarray = image, transform, surfarray
    for name, mod in {"pygame": pg, "pygame.image": image, "pygame.transform": transform,
                      "pygame.surfarray": surfarray}.items():
        modules[name] = mod
    return pg
