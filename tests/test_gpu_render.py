"""GPU: rgb_array rendering — the CUDA rasteriser (csrc/pz_render.cu) against a numpy
rasteriser of the same display lists (the lists themselves are pinned against the reference's draw() on CPU,
tests/test_render.py), the batch renderer on selected envs of a large batch, and the facade's render_mode="rgb_array".
Sprites: the reference's own PNGs when oracle/stage_ref.py staged them (oracle/_ref, git-ignored), random
binary-alpha sprites of the right sizes otherwise. Rasterisation is unpinned against pygame (not installable)."""

import json
import os

import numpy as np
import pytest
import torch

from pikazoo_b200 import render as R
from tests.conftest import ROOT
from tests.test_render import sizes_from_golden

pytestmark = pytest.mark.gpu

STAGED = os.path.join(ROOT, "oracle", "_ref", "pikazoo", "env", "img")


@pytest.fixture(scope="module")
def sprites():
    if os.path.isdir(STAGED):
        return R.SpriteSet(STAGED)
    with open(os.path.join(ROOT, "tests", "golden", "render.json")) as f:
        size = sizes_from_golden(json.load(f))
    rng = np.random.default_rng(1)
    images = {}
    for name in R.STATIC_FILES + R.DYNAMIC_FILES:
        w, h = size[name]
        im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        im[..., 3] = np.where(rng.random((h, w)) < 0.7, 255, 0)
        if name in R.STATIC_FILES[:1]:
            im[..., 3] = 255
        images[name] = im
    return R.SpriteSet(images=images)


def test_real_sprites_have_the_sizes_the_reference_saw(sprites):
    with open(os.path.join(ROOT, "tests", "golden", "render.json")) as f:
        size = sizes_from_golden(json.load(f))
    for name, wh in sprites.size.items():
        assert tuple(size[name]) == tuple(wh), name
    assert sprites.background.shape == (304, 432, 3) and len(sprites.background_items) == 446


def test_cuda_rasteriser_equals_numpy_rasteriser(cuda_lib, sprites):
    import pikazoo_b200

    n = 4096
    env = pikazoo_b200.PikaVecEnv(n, seed=3, winning_score=12, serve="random", is_player2_computer=True)
    idx = [0, 5, 777, 4095]
    rend = env.attach_renderer(idx, sprites=sprites, cloud_seed=7)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    seen = set()
    for t in range(400):
        env.step(torch.randint(0, 18, (n, 2), generator=g, device="cuda", dtype=torch.int32))
        frames = rend.render()
        assert frames.shape == (4, 304, 432, 3) and frames.dtype == torch.uint8
        if t % 20 == 0 or any(it[0] in ("ball_punch.png", "ball_trail.png") for lst in rend.last_items for it in lst):
            got = frames.cpu().numpy()
            for b, lst in enumerate(rend.last_items):
                want = R.composite(lst, sprites, sprites.background.copy())
                assert np.array_equal(got[b], want), (t, b)
                seen |= {it[0] for it in lst}
    assert {"ball_punch.png", "cloud.png", "wave.png", "shadow.png"} <= seen
    # the sky row above the clouds' range is pure background; frames of different envs differ
    assert not np.array_equal(got[0], got[1])
    env.detach_renderer()
    env.rollout(8, actions="synth")


def test_rendering_does_not_touch_the_game(cuda_lib, sprites):
    """With a render mode the reference's clouds draw from the game's generator; here they have their own: the same
    seed gives the same game with and without rendering."""
    from pikazoo_b200 import pikazoo_v0

    a = pikazoo_v0.env(winning_score=2, serve="random", seed=5, is_player2_computer=True)
    b = pikazoo_v0.env(winning_score=2, serve="random", seed=5, is_player2_computer=True, render_mode="rgb_array",
                       sprite_dir=STAGED if os.path.isdir(STAGED) else None) if os.path.isdir(STAGED) else None
    if b is None:
        pytest.skip("the facade loads sprites from a directory; none staged here")
    oa, _ = a.reset()
    ob, _ = b.reset()
    f0 = b.render()
    assert f0.shape == (304, 432, 3) and f0.dtype == np.uint8 and a.render() is None
    rng = np.random.default_rng(0)
    changed = steps = 0
    for t in range(300):
        if not a.agents:
            break
        act = {"player_1": int(rng.integers(0, 18)), "player_2": 0}
        ra, rb = a.step(act), b.step(act)
        assert np.array_equal(ra[0]["player_1"], rb[0]["player_1"]) and ra[1] == rb[1] and ra[2] == rb[2]
        f = b.render()
        steps += 1
        changed += int(not np.array_equal(f, f0))
        f0 = f
    assert steps > 50 and changed == steps  # the scene animates (clouds and waves move every rendered frame)
    assert np.array_equal(a.state_words(), b.state_words())
    # player 1's sprite is where the state says: the 64x64 box around (x, y) differs from the background there
    st = b.state_words()
    x, y = int(st[0]), int(st[1])
    box = f0[max(y - 32, 0):y + 32, x - 32:x + 32]
    bg = b._sprites.background[max(y - 32, 0):y + 32, x - 32:x + 32]
    assert not np.array_equal(box, bg)
