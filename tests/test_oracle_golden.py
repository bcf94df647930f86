"""CPU: the C oracle reproduces every golden session recorded from the unmodified reference
(tests/golden/sessions.json, 1,039 full games; generator: oracle/make_golden.py)."""

import pytest

from tests.helpers import OracleStepper, replay_group


def test_golden_inventory(golden):
    assert golden["total_games"] >= 1000
    assert all(g["oracle_checked"] for g in golden["groups"])
    names = {g["name"] for g in golden["groups"]}
    assert {"ai_vs_ai_ws15_winner", "random_ws15_winner", "simplify_shaped_ws15",
            "random_ws5_serve_random"} <= names


def test_survey_known_answer(golden):
    # config 1, seed 0 -> 13,987 frames, 15-5
    g = next(g for g in golden["groups"] if g["name"] == "ai_vs_ai_ws15_winner")
    assert g["sessions"][0]["episodes"] == [{"frames": 13987, "scores": [15, 5]}]
    assert g["sessions"][1]["episodes"] == [{"frames": 11383, "scores": [15, 4]}]
    # seed 2 never terminates (frame cap)
    assert g["sessions"][2]["episodes"] == []


@pytest.mark.parametrize("name", [
    "ai_vs_ai_ws15_winner", "random_ws15_winner", "simplify_shaped_ws15", "random_ws5_serve_random",
    "ai_p1_vs_random_ws7_alternate", "random_vs_ai_p2_ws7_random", "ai_vs_ai_ws3_random_multi",
])
def test_oracle_replays_golden_group(golden, name):
    group = next(g for g in golden["groups"] if g["name"] == name)
    if name == "ai_vs_ai_ws15_winner":
        # 1.5 M frames of per-env hashing in Python is slow: first 12 sessions on CPU (all 100 on the GPU suite)
        group = dict(group, num_envs=12, sessions=group["sessions"][:12])
    bad = replay_group(group, OracleStepper)
    assert not bad, bad[:5]


WRAPPER_GROUPS = ["normalize_rins_outer_shaped_simplify_ws5", "rins_inner_shaped_record_ws5",
                  "normalize_record_ai_p2_ws3", "normalize_rins_int_ai_vs_ai_ws2"]


@pytest.mark.parametrize("name", WRAPPER_GROUPS)
def test_oracle_replays_wrapper_stacks(golden_wrappers, name):
    """NormalizeObservation / RewardInNormalState (inside and outside RewardByBallPosition) /
    RecordEpisodeStatistics as recorded from the reference's own wrapper classes."""
    group = next(g for g in golden_wrappers["groups"] if g["name"] == name)
    assert group["oracle_checked"]
    bad = replay_group(group, OracleStepper)
    assert not bad, bad[:5]


def test_oracle_reproduces_survey_known_answers():
    """frames / final scores / trajectory hashes the survey probed on the unmodified reference
    (SURVEY.md §8(c)), e.g. config 1 seed 0 -> 13,987 frames, 15-5, 7c7cc240a767c583"""
    from tests.helpers import check_survey_known_answers

    check_survey_known_answers(lambda n, seed, cfg: OracleStepper(n, seed, cfg))
