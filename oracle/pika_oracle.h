/* TEST INFRASTRUCTURE — CPU restatement of the reference hot path (oracle).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (pikazoo_b200/) never does.
 *
 * Parity status: the reference holds no golden vectors or known-answer tests for this
 * path (SURVEY.md §4, §8(c) — "parity unpinned" at the reference's own test level), so
 * this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build
 * container: the fixtures in tests/golden/ are produced by oracle/make_golden.py from the
 * unmodified /root/reference code (via oracle/ref_harness.py) and tests/test_oracle_*.py
 * check this file against them and, when /root/reference is present, against the live
 * reference frame by frame.
 *
 * Follows: pikazoo/env/physics.py:59-99,181-218,258-277,280-884,
 *          pikazoo/env/pikazoo_env.py:119-141,149-248,576-624,
 *          pikazoo/wrappers/simplify_action.py:16-25,
 *          pikazoo/wrappers/reward_by_ball_position.py:20-31,
 *          pikazoo/wrappers/normalize_observation.py:18-32,
 *          pikazoo/wrappers/reward_in_normal_state.py:10-15,
 *          pikazoo/wrappers/record_episode_statistics.py:17-40,
 *          numpy 2.3.5 (unpinned in the reference: pyproject.toml:25 "numpy>=1.21.0"):
 *          SeedSequence, PCG64 (XSL-RR 128/64), Generator.integers scalar path
 *          (buffered 32-bit Lemire), as restated in SURVEY.md §8(c).
 */
#ifndef PIKA_ORACLE_H
#define PIKA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 13 words, physics.py:140-218 + PikaUserInput.power_hit_key_is_down_previous (physics.py:51) */
typedef struct {
    int32_t x, y, y_velocity, state, frame_number, delay_before_next_frame;
    int32_t normal_status_arm_swing_direction, diving_direction, lying_down_duration_left;
    int32_t is_collision_with_ball_happened, computer_boldness, computer_where_to_stand_by;
    int32_t power_hit_key_is_down_previous;
} pk_player;

/* 11 words, physics.py:221-277 (render-only fields omitted) */
typedef struct {
    int32_t x, y, x_velocity, y_velocity, previous_x, previous_y;
    int32_t previous_previous_x, previous_previous_y, is_power_hit;
    int32_t expected_landing_point_x, punch_effect_x;
} pk_ball;

/* 53 int32 words: the "unpacked parity state" also produced by pz_export_state. */
typedef struct {
    pk_player p[2];
    pk_ball b;
    int32_t scores[2], round_ended, game_ended, is_player2_serve; /* pikazoo_env.py:100-111 */
    uint32_t rng_state[4], rng_inc[4]; /* PCG64 128-bit state / inc, little-endian words */
    uint32_t has_uint32, uinteger;     /* numpy's buffered high half of next64 */
    /* word 52, not reference env state: step() calls since the last reset(), i.e. what
     * RecordEpisodeStatistics.episode_lengths would hold (record_episode_statistics.py:33) */
    int32_t episode_frames;
} pk_env;

enum { PK_SERVE_WINNER = 0, PK_SERVE_ALTERNATE = 1, PK_SERVE_RANDOM = 2 };

typedef struct {
    int32_t winning_score;        /* pikazoo_env.py:102 */
    int32_t serve;                /* PK_SERVE_*, pikazoo_env.py:104-105 */
    int32_t is_player1_computer;  /* pikazoo_env.py:97 */
    int32_t is_player2_computer;
    int32_t simplify_action;      /* SimplifyAction wrapper on: actions in [0,13) */
    int32_t reward_by_ball_position; /* RewardByBallPosition wrapper on */
    int32_t x_line, y_line;       /* reward_by_ball_position.py:11-12 */
    double additional_reward[8];  /* reward_by_ball_position.py:10 */
    /* RewardInNormalState(env, reward): 0 off, 1 = outermost (applied after RewardByBallPosition),
     * 2 = innermost (applied before it); reward_in_normal_state.py:10-15 */
    int32_t reward_in_normal_state;
    /* product-defined truncation (the reference never truncates): 0 off */
    int32_t max_episode_frames;
    double normal_state_reward;
} pk_config;

int pk_env_words(void); /* sizeof(pk_env)/4 == 53 */

/* numpy: PCG64(SeedSequence(seed)) -> rng_state/rng_inc, has_uint32 = uinteger = 0. */
void pk_pcg64_seed(uint64_t seed, uint32_t state[4], uint32_t inc[4]);
/* numpy: Generator.integers(0, high) scalar path on the env's stream. */
int32_t pk_integers(pk_env *e, uint32_t high);

/* A freshly constructed reference env object whose generator was then overwritten with
 * PCG64(seed).state (protocol S0): constructor defaults + seeded stream. reset() has NOT
 * been called yet. */
void pk_init(pk_env *e, uint64_t seed);

/* raw_env.reset (pikazoo_env.py:149-173) on a live object; obs = [obs_p1(35), obs_p2(35)]. */
void pk_reset(pk_env *e, const pk_config *c, int32_t obs[70]);

/* (wrapped) env.step. Returns 0, or -1 if an action is out of range (reference raises).
 * reward[2] as Python would compute it (int, or int + float in double). */
int pk_step(pk_env *e, const pk_config *c, int32_t a1, int32_t a2, int32_t obs[70],
            double reward[2], uint8_t *terminated);

/* Batched NEXT-STEP auto-reset semantics of the product: for
 * each env, if game_ended at call time: autoreset ? reset() (obs = reset obs, reward 0,
 * done 0, action ignored) : no-op (obs re-emitted, reward 0, done 1); else step().
 * actions int32 [n][2]; obs int32 [n][2][35]; reward double [n][2]; done uint8 [n]. */
int pk_vec_step(pk_env *envs, int64_t n, const pk_config *c, const int32_t *actions,
                int32_t *obs, double *reward, uint8_t *done, int autoreset);

void pk_vec_obs(const pk_env *envs, int64_t n, int32_t *obs); /* _get_obs of every env, no step */

/* pk_vec_step plus RecordEpisodeStatistics (record_episode_statistics.py:17-40) and the product's
 * truncation. ep_return double [n][2] in/out: running sum of each agent's wrapped rewards, zeroed by
 * a reset (:24-25), += on every step (:32); ep_length int32 [n]: step() calls since reset (:33);
 * truncated uint8 [n]: the episode reached c->max_episode_frames without terminating. A truncated
 * env is treated like a terminated one by the next call (reset, or frozen). Any pointer may be NULL. */
int pk_vec_step_ex(pk_env *envs, int64_t n, const pk_config *c, const int32_t *actions, int32_t *obs,
                   double *reward, uint8_t *done, int autoreset, double *ep_return,
                   int32_t *ep_length, uint8_t *truncated);
void pk_vec_reset_ex(pk_env *envs, int64_t n, const pk_config *c, int32_t *obs, double *ep_return,
                     int32_t *ep_length);

/* NormalizeObservation (normalize_observation.py:18-32): (obs - low) / (high - low) in double with the
 * bounds of raw_env.observation_space (pikazoo_env.py:485-562); obs int32 [n][70] -> out double [n][70]. */
void pk_normalize_obs(int64_t n, const int32_t *obs, double *out);
void pk_vec_init(pk_env *envs, int64_t n, uint64_t base_seed);
void pk_vec_reset(pk_env *envs, int64_t n, const pk_config *c, int32_t *obs);

/* The product's on-device synthetic action stream for rollouts (NOT part of the
 * reference; defined in DESIGN.md): uniform in [0, n_actions). */
int32_t pk_synth_action(uint64_t action_seed, uint64_t global_env, uint64_t frame, int agent,
                        uint32_t n_actions);

/* K frames of every env with auto-reset; action_mode 0 = all actions 0 (NOOP; computer
 * players ignore them), 1 = pk_synth_action(action_seed, first_env + i, frame0 + k, agent).
 * stats[9] += {env_steps, episodes, sum_episode_frames, p1_wins, p2_wins,
 *              p1_points, p2_points, resets, truncated}. Returns total env-steps executed. */
int64_t pk_vec_rollout(pk_env *envs, int64_t n, const pk_config *c, int K, int action_mode,
                       uint64_t action_seed, uint64_t first_env, uint64_t frame0,
                       int64_t *stats);

/* Test helper: calculate_expected_landing_point_x_for (power = 0, physics.py:643-686) or the loop of
 * expected_landing_point_x_when_power_hit (power = 1, physics.py:847-884) started from
 * xyv[i] = {x, y, x_velocity, y_velocity}; out[i] = landing x. */
void pk_simulate_many(int64_t n, const int32_t *xyv, int power, int32_t *out);

#ifdef __cplusplus
}
#endif
#endif
