"""CPU: the oracle's numpy-RNG restatement against numpy itself (the third-party dependency
that actually runs under the reference: SeedSequence -> PCG64 -> Generator.integers)."""

import numpy as np
import pytest

from oracle import pyoracle as po


def _u128(words):
    return sum(int(words[k]) << (32 * k) for k in range(4))


def test_pcg64_seed_known_answer():
    # SURVEY.md §8(c) KAT
    st, inc = po.pcg64_seed(0)
    assert _u128(st) == 0x1AA1B5345996452D09585EB7A69561E3
    assert _u128(inc) == 0x418DDADB3AF71A82588133BC447873A9


@pytest.mark.parametrize("seed", [0, 1, 2, 7, 12345, 2**31, 2**32 - 1, 2**32, 2**40 + 7, 2**63, 2**64 - 1])
def test_pcg64_seed_matches_numpy(seed):
    st, inc = po.pcg64_seed(seed)
    ref = np.random.PCG64(seed).state["state"]
    assert _u128(st) == ref["state"] and _u128(inc) == ref["inc"]


def test_integers_known_answer():
    env = np.zeros(po.ENV_WORDS, dtype=np.int32)
    po.lib().pk_init(po._p(env), 0)
    got = [po.lib().pk_integers(po._p(env), 5) for _ in range(10)]
    assert got == [4, 3, 2, 1, 1, 0, 0, 0, 0, 4]


@pytest.mark.parametrize("seed", [0, 3, 99, 2**33 + 1])
def test_integers_stream_matches_numpy(seed):
    env = np.zeros(po.ENV_WORDS, dtype=np.int32)
    po.lib().pk_init(po._p(env), seed)
    g = np.random.Generator(np.random.PCG64(seed))
    pick = np.random.default_rng(seed + 1)
    for _ in range(5000):
        high = int(pick.choice([2, 3, 5, 20]))
        assert po.lib().pk_integers(po._p(env), high) == int(g.integers(0, high))
    # internal state too (buffered half included)
    st = g.bit_generator.state
    assert _u128(env.view(np.uint32)[42:46]) == st["state"]["state"]
    assert int(env.view(np.uint32)[50]) == st["has_uint32"]
    if st["has_uint32"]:
        assert int(env.view(np.uint32)[51]) == st["uinteger"]


def test_synth_action_python_matches_c():
    from oracle.synth import synth_action, synth_actions_numpy

    for (s, e, f, a, n) in [(0, 0, 0, 0, 18), (0x5EED, 5, 77, 1, 13), (2**63 + 5, 2**40, 2**33 + 1, 1, 18)]:
        assert synth_action(s, e, f, a, n) == po.synth_action(s, e, f, a, n)
    v = synth_actions_numpy(0x5EED, 10, 64, 1234, 18)
    for i in range(64):
        for a in (0, 1):
            assert v[i, a] == po.synth_action(0x5EED, 10 + i, 1234, a, 18)
    assert v.min() >= 0 and v.max() < 18
