"""GPU: the PLAIN instantiations of the per-step kernels (pz_device.cuh: the options compiled out for a plain batched
run) against the general instantiations of the same kernels. Asking for the status byte is enough to make a launch
take the general kernel (pz_step_inst.inc:launch_dt), and it changes nothing else: two envs built alike, one of them
with status=True, fed the same actions, must produce the same observations (as bit patterns), rewards, done flags,
statistics and packed state on every step — for every computer-player mask, row dtype, layout and action dtype. The
general kernels are the ones the configuration fuzz (always with episode statistics) runs against the oracle."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_ROWS = [(torch.int32, False), (torch.int16, False), (torch.float32, True), (torch.float16, True),
         (torch.bfloat16, True), (torch.float64, True)]


def _bits(t: torch.Tensor) -> torch.Tensor:
    return t.view(torch.int16) if t.dtype == torch.bfloat16 else t


@pytest.mark.parametrize("mask", [0, 1, 2, 3])
@pytest.mark.parametrize("layout", ["env_major", "feature_major"])
@pytest.mark.parametrize("row", range(len(_ROWS)))
def test_plain_kernel_equals_general_kernel(cuda_lib, mask, layout, row):
    import pikazoo_b200

    obs_dtype, normalize = _ROWS[row]
    act_dtype = [torch.int32, torch.uint8, torch.int64][(mask + row) % 3]
    n = 4096 + 45  # ragged last warp and CTA
    kw = dict(seed=1234 + 17 * mask + row, winning_score=2, serve=["winner", "random", "alternate"][row % 3],
              is_player1_computer=bool(mask & 1), is_player2_computer=bool(mask & 2), obs_dtype=obs_dtype,
              normalize_observation=normalize, obs_layout=layout, action_dtype=act_dtype)
    if layout == "feature_major":
        kw["obs_feature_rows"] = 40
    plain = pikazoo_b200.PikaVecEnv(n, **kw)
    general = pikazoo_b200.PikaVecEnv(n, status=True, **kw)
    o1, o2 = plain.reset(), general.reset()
    assert torch.equal(_bits(o1), _bits(o2))
    g = torch.Generator(device="cuda").manual_seed(99 + mask)
    for step in range(260):
        a = torch.randint(0, 18, (n, 2), generator=g, device="cuda").to(act_dtype)
        r1 = plain.step(a)
        r2 = general.step(a)
        if step % 13 == 0 or step > 250:
            assert torch.equal(_bits(r1[0][..., :n] if layout == "feature_major" else r1[0]),
                               _bits(r2[0][..., :n] if layout == "feature_major" else r2[0])), step
            assert torch.equal(r1[1], r2[1]) and torch.equal(r1[2], r2[2]), step
    assert torch.equal(plain.state, general.state)
    s1, s2 = plain.stats_dict(), general.stats_dict()
    assert s1 == s2
    if mask != 3:  # (two computer players do not finish a game in 260 frames)
        assert s1["episodes"] > 0 and s1["resets"] > 0
