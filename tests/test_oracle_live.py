"""CPU, build container only: the C oracle against the LIVE unmodified reference, frame by
frame including the full hidden state. Skipped where /root/reference is absent (GPU box)."""

import pytest

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


def test_reference_known_answers():
    # SURVEY.md §8(c) hashes (sha256 over obs + reward byte), config 1 seed 0 and random-vs-random seed 0
    r = rh.play_game(0, lambda f: (0, 0), is_player1_computer=True, is_player2_computer=True,
                     winning_score=15, serve="winner")
    assert (r["frames"], r["scores"], r["hash16"]) == (13987, [15, 5], "7c7cc240a767c583")


@pytest.mark.parametrize("cfg,mode", [
    (dict(is_player1_computer=True, is_player2_computer=True, winning_score=2, serve="random"), "noop"),
    (dict(winning_score=3, serve="alternate"), "synth"),
    (dict(winning_score=3, serve="winner", simplify_action=True,
          reward_by_ball_position=((1, 2, 3, 4, 5, 6, 7, 8), 200, 150)), "synth"),  # int rewards, moved lines
    (dict(is_player2_computer=True, winning_score=2, serve="winner", simplify_action=True), "synth"),
])
def test_oracle_matches_live_reference(cfg, mode):
    from oracle.make_golden import run_session  # asserts obs / reward / term / 52-word state every frame

    for seed in (11, 12, 13):
        out = run_session((cfg, mode, seed, seed, 2, 20000))
        assert out["calls"] > 100


@pytest.mark.parametrize("cfg,mode", [
    (dict(winning_score=2, serve="random", reward_by_ball_position=((0.5, 0, -0.5, 0.25, 0, 0.125, 0, -1), 216, 176),
          reward_in_normal_state=0.25, normalize_observation=True, record_episode_statistics=True), "synth"),
    (dict(winning_score=2, serve="winner", is_player1_computer=True, reward_in_normal_state=-1, normal_state_first=True,
          reward_by_ball_position=((1, 0, -1, 0, 0, 1, 0, -1), 216, 176), record_episode_statistics=True), "synth"),
])
def test_oracle_wrappers_match_live_reference(cfg, mode):
    """the reference's NormalizeObservation / RewardInNormalState / RecordEpisodeStatistics classes, live"""
    from oracle.make_golden import run_session

    for seed in (21, 22):
        out = run_session((cfg, mode, seed, seed, 2, 20000))
        assert out["calls"] > 100 and all("returns" in e for e in out["episodes"])
