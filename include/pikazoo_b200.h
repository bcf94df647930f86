/* pikazoo_b200 — C ABI of the B200-native batched Pikachu-Volleyball simulator.
 *
 * This is the drop-in boundary for the reference's hot path. The reference
 * (helpingstar/pika-zoo) is pure Python and has no FFI of its own; the interface replaced
 * is the PettingZoo ParallelEnv surface of `raw_env`, so each entry point below cites the
 * reference method whose work it performs for a whole batch of independent envs:
 *
 *   pz_seed / pz_seed_array   raw_env.__init__ + _seed            pikazoo/env/pikazoo_env.py:79-147,570-571
 *                             (+ protocol S0 generator overwrite)
 *   pz_reset                  raw_env.reset                       pikazoo/env/pikazoo_env.py:149-173
 *   pz_step                   raw_env.step                        pikazoo/env/pikazoo_env.py:175-240
 *                             -> physics_engine and callees       pikazoo/env/physics.py:59-99,280-884
 *                             + SimplifyAction.step               pikazoo/wrappers/simplify_action.py:16-25
 *                             + RewardByBallPosition.step         pikazoo/wrappers/reward_by_ball_position.py:20-31
 *                             + raw_env._get_obs                  pikazoo/env/pikazoo_env.py:576-624
 *                             + NormalizeObservation.step/reset   pikazoo/wrappers/normalize_observation.py:18-32
 *                             + RewardInNormalState.step          pikazoo/wrappers/reward_in_normal_state.py:10-15
 *   pz_step_ex                pz_step + RecordEpisodeStatistics   pikazoo/wrappers/record_episode_statistics.py:17-40
 *                             (per-env running return / length) and optional truncation
 *   pz_rollout                K x raw_env.step with the state held in registers
 *   pz_host_step              raw_env.step for callers holding HOST buffers (numpy users)
 *   pz_export_state / pz_import_state   the Python object graph <-> packed device state
 *
 * Conventions: every pointer named *_dev is device memory owned by the caller (the library
 * allocates nothing except inside pz_host_ctx); `stream` is a cudaStream_t passed as void*;
 * all calls are asynchronous on that stream and re-entrant; return value 0 = success,
 * > 0 = cudaError_t, < 0 = PZ_E_*; nothing throws across the boundary.
 * Binary is sm_100a only. There is no CPU fallback.
 */
#ifndef PIKAZOO_B200_H
#define PIKAZOO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZ_VERSION 3

/* packed device state: int32 words per env (structure-of-arrays) */
#define PZ_STATE_WORDS 17
/* unpacked parity state: int32 words per env (layout of oracle/pika_oracle.h pk_env) */
#define PZ_UNPACKED_WORDS 53
#define PZ_OBS_WORDS 35 /* per agent, pikazoo_env.py:481-565 */
#define PZ_NUM_STATS 16

enum { PZ_SERVE_WINNER = 0, PZ_SERVE_ALTERNATE = 1, PZ_SERVE_RANDOM = 2 }; /* pikazoo_env.py:104 */
enum { PZ_ACT_I32 = 0, PZ_ACT_I64 = 1, PZ_ACT_U8 = 2 };
enum { PZ_REW_F32 = 0, PZ_REW_F64 = 1 };
/* element type of the observation rows [n][2][35]. I32 is the reference's declared dtype
 * (pikazoo_env.py:564); I16 is the same integers in half the bytes (every value fits: |v| <= 32767
 * is enforced by saturation, play stays within +-700); the float types hold (float)value, or the
 * NormalizeObservation output when pz_config.normalize_observation is set. F64 is what the reference
 * wrapper itself produces (numpy true division of int64 arrays); F32 = float32(F64 value) exactly;
 * F16 / BF16 = round-to-nearest-even of the F32 value. */
enum { PZ_OBS_I32 = 0, PZ_OBS_I16 = 1, PZ_OBS_F32 = 2, PZ_OBS_F16 = 3, PZ_OBS_BF16 = 4, PZ_OBS_F64 = 5 };
/* memory layout of obs_dev.
 *   ENV_MAJOR     [n][2][35]: one row per env and agent, the reference's layout (obs[i][a] = env i's
 *                 observation for player a+1);
 *   FEATURE_MAJOR [2][obs_feature_rows][n]: element k of agent a's observation of env i at
 *                 ((a * obs_feature_rows) + k) * n + i — the layout a device-side policy wants (every
 *                 feature a contiguous vector over envs: GEMM operands with leading dimension n, no
 *                 transposing pass). Rows >= 35 (padding up to a multiple of 8 for tensor-core GEMMs) are
 *                 never written; zero them once. Device pointers only (not pz_host_*).
 *   ENV_MAJOR_SHARED [n][35], PZ_OBS_I32 / PZ_OBS_I16 only: player_1's row alone. player_2's observation holds the
 *                 same 35 values with the two player blocks swapped (pikazoo_env.py:585-586):
 *                 obs_p2[k] = row[pz_obs_player2_index(k)] = row[k + 13] (k < 13), row[k - 13] (13 <= k < 26), row[k].
 *                 As int16 this is 70 B per env instead of the 280 B of int32 [n][2][35] with nothing lost — the
 *                 format for callers behind PCIe (pz_host_*). */
enum { PZ_LAYOUT_ENV_MAJOR = 0, PZ_LAYOUT_FEATURE_MAJOR = 1, PZ_LAYOUT_ENV_MAJOR_SHARED = 2 };
int pz_obs_player2_index(int k); /* -1 outside [0, 35) */
/* RewardInNormalState composition order relative to RewardByBallPosition */
enum { PZ_RINS_OFF = 0, PZ_RINS_OUTER = 1 /* RewardInNormalState(RewardByBallPosition(env)) */,
       PZ_RINS_INNER = 2 /* RewardByBallPosition(RewardInNormalState(env)) */ };
enum { PZ_ACTIONS_NOOP = 0, PZ_ACTIONS_SYNTH = 1 }; /* pz_rollout action source */

enum {
    PZ_E_BADARG = -1,    /* null pointer, n < 0, K < 1 ... */
    PZ_E_BADCONFIG = -2, /* winning_score outside [1,1023], unknown serve/dtype code */
    PZ_E_ALIGN = -3,     /* state/obs pointer not 16-byte aligned */
    PZ_E_NODEVICE = -4,  /* no sm_100 device / kernel image unusable */
    PZ_E_ABI = -5        /* pz_config.struct_bytes / abi_version do not match this library: the caller's binding was
                            written against another revision of this header (see pz_config_init) */
};

/* indices into the int64 statistics vector (all counters are += over calls) */
enum {
    PZ_STAT_CALLS = 0,       /* env calls issued (n per pz_step, n*K per pz_rollout);
                                raw_env.step calls executed = CALLS - RESETS - FROZEN */
    PZ_STAT_EPISODES = 1,    /* games terminated */
    PZ_STAT_EPISODE_FRAMES = 2, /* sum of episode lengths of terminated games */
    PZ_STAT_P1_WINS = 3,
    PZ_STAT_P2_WINS = 4,
    PZ_STAT_P1_POINTS = 5,   /* points of terminated games (sum of final scores) */
    PZ_STAT_P2_POINTS = 6,
    PZ_STAT_RESETS = 7,      /* raw_env.reset calls executed by auto-reset */
    PZ_STAT_BAD_ACTIONS = 8, /* actions outside the action space (treated as action 0) */
    PZ_STAT_FROZEN = 9,      /* calls on terminated envs with autoreset off (no-ops) */
    PZ_STAT_TRUNCATED = 10   /* episodes cut by max_episode_frames */
};

/* Passed by pointer to every entry point that takes a configuration. The first two members are the ABI
 * handshake: pz_config_init() fills them, and every entry point returns PZ_E_ABI — before reading anything
 * else — unless struct_bytes == pz_config_bytes() and abi_version == pz_version(). A binding in another
 * language (ctypes, cffi, cgo ...) that declares a stale, shorter or longer struct is therefore refused
 * instead of being read past its end. New members are only ever appended, with a PZ_VERSION bump. */
typedef struct pz_config {
    uint32_t struct_bytes;           /* sizeof(pz_config) as the caller's binding declares it */
    uint32_t abi_version;            /* PZ_VERSION the caller's binding was written against */
    int32_t winning_score;           /* pikazoo_env.py:81,102; 1..1023 */
    int32_t serve;                   /* PZ_SERVE_*; pikazoo_env.py:82,104-105 */
    int32_t is_player1_computer;     /* pikazoo_env.py:83 */
    int32_t is_player2_computer;     /* pikazoo_env.py:84 */
    int32_t simplify_action;         /* fuse SimplifyAction: actions in [0,13) */
    int32_t reward_by_ball_position; /* fuse RewardByBallPosition */
    int32_t x_line, y_line;          /* reward_by_ball_position.py:11-12 (defaults 216, 176) */
    double additional_reward[8];     /* reward_by_ball_position.py:10 */
    int32_t autoreset;               /* 1: a call on a terminated env performs reset() (NEXT-STEP) */
    int32_t action_dtype;            /* PZ_ACT_*: element type of actions_dev [n][2] */
    int32_t reward_dtype;            /* PZ_REW_*: element type of reward_dev [n][2] */
    int32_t flags;                   /* PZ_FLAG_* */
    int32_t obs_dtype;               /* PZ_OBS_*: element type of obs_dev */
    int32_t normalize_observation;   /* fuse NormalizeObservation: (obs - low) / (high - low) with the bounds of
                                        pikazoo_env.py:485-562; float obs dtypes only */
    int32_t reward_in_normal_state;  /* PZ_RINS_*: fuse RewardInNormalState (reward_in_normal_state.py:10-15) */
    int32_t max_episode_frames;      /* 0: never truncate (the reference never does). > 0: an env whose episode
                                        reaches this many step() calls without terminating is truncated: it
                                        reports truncated = 1 on that call and is reset (or frozen) by the next */
    double normal_state_reward;      /* RewardInNormalState(env, reward) */
    int32_t obs_layout;              /* PZ_LAYOUT_* */
    int32_t obs_feature_rows;        /* FEATURE_MAJOR: rows per agent, >= 35 (0 means 35) */
} pz_config;

/* pz_config.flags */
#define PZ_FLAG_NO_TABLES 1 /* computer players: always run the trajectory simulations iteratively
                               instead of reading the memoised landing tables (results are identical) */

#define PZ_FLAG_NO_L2_HINTS 2 /* do not mark the output stores evict-first in L2; for A/B
                                 measurements */

#define PZ_FLAG_NO_PDL 4 /* launch pz_step without programmatic stream serialization; for A/B
                            measurements */

int pz_version(void);
int pz_state_words(void);
int pz_unpacked_words(void);
size_t pz_state_bytes(int64_t n);
const char *pz_strerror(int code);
size_t pz_config_bytes(void); /* sizeof(pz_config) in this library */
/* Reference defaults (ws=15, serve winner, no computer players, auto-reset on) into *cfg, and the handshake
 * members. caller_struct_bytes is sizeof() of the CALLER's declaration of pz_config (C: sizeof(pz_config);
 * ctypes: ctypes.sizeof(Cfg)): if it differs from pz_config_bytes() nothing is written and PZ_E_ABI returned. */
int pz_config_init(pz_config *cfg, size_t caller_struct_bytes);

/* Fresh env objects, generator of env i = numpy PCG64(SeedSequence(base_seed + first_env + i)) — protocol S0 of the
 * parity tests, and what makes trajectories independent of how a batch is sharded (first_env = the shard's first
 * global index). Consecutive base seeds therefore give SHIFTED, almost identical batches (env i of seed s + 1 is env
 * i + 1 of seed s): independent replicates must use base seeds at least the total batch size apart, or explicit
 * per-env seeds (pz_seed_array). reset() has not been called. */
int pz_seed(int32_t *state_dev, int64_t n, uint64_t base_seed, uint64_t first_env, void *stream);
/* Same with explicit per-env seeds (device array of n uint64). */
int pz_seed_array(int32_t *state_dev, int64_t n, const uint64_t *seeds_dev, void *stream);

/* reset() on every env. obs_dev: [n][2][35] of cfg->obs_dtype, or NULL. */
int pz_reset(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, void *stream);

/* raw_env._get_obs (pikazoo_env.py:576-624) of every env from the state as it stands (no frame is run):
 * obs_dev as for pz_reset. */
int pz_observe(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, void *stream);

/* One frame of every env.
 *   actions_dev [n][2] (cfg->action_dtype), obs_dev [n][2][35] of cfg->obs_dtype (may be NULL),
 *   reward_dev [n][2] (cfg->reward_dtype, may be NULL), done_dev uint8 [n] (may be NULL),
 *   stats_dev int64 [PZ_NUM_STATS] (may be NULL).
 * Env terminated before the call: autoreset ? reset() (reward 0, done 0) : frozen (done 1). */
int pz_step(int32_t *state_dev, int64_t n, const pz_config *cfg, const void *actions_dev,
            void *obs_dev, void *reward_dev, uint8_t *done_dev, int64_t *stats_dev, void *stream);

/* pz_step plus per-env episode bookkeeping (RecordEpisodeStatistics, record_episode_statistics.py:17-40)
 * and truncation. Every member may be NULL.
 *   episode_return_dev  double [n][2], in/out, caller-owned like the state: the running sum of each
 *                       agent's (wrapped) rewards since the last reset, in step order as Python adds them
 *                       (record_episode_statistics.py:32); zeroed by a reset, so on the call that ends an
 *                       episode it holds infos[agent]["episode"]["r"].
 *   episode_length_dev  int32 [n], out: step() calls since the last reset (["episode"]["l"] when done).
 *   truncated_dev       uint8 [n], out: 1 on the call where the episode reached cfg->max_episode_frames
 *                       without terminating (and on every later call while frozen). */
typedef struct pz_episode_io {
    double *episode_return_dev;
    int32_t *episode_length_dev;
    uint8_t *truncated_dev;
    uint8_t *status_dev; /* uint8 [n], out: (player_1's BASE reward + 1) | terminated << 2 | truncated << 3 — reward,
                            done and truncation of an unshaped env in one byte (player_2's reward is the negative);
                            on reset / frozen calls the reward field is 1 (reward 0) */
    uint32_t *seq_dev;   /* optional, batches of at most 128 envs (one CTA): after every output of the call is
                            written, seq_value is stored here behind a system-scope fence. With host-mapped buffers
                            (cudaHostAlloc) a host thread can spin on this word instead of synchronising the stream:
                            the single-env facade's completion signal */
    uint32_t seq_value;
} pz_episode_io;
int pz_reset_ex(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, const pz_episode_io *episode,
                void *stream); /* pz_reset that also zeroes episode_return_dev / episode_length_dev */
int pz_step_ex(int32_t *state_dev, int64_t n, const pz_config *cfg, const void *actions_dev, void *obs_dev,
               void *reward_dev, uint8_t *done_dev, int64_t *stats_dev, const pz_episode_io *episode,
               void *stream);

/* K frames of every env in one launch, state register-resident, auto-reset always on.
 * action_source PZ_ACTIONS_NOOP: both actions 0 (computer players decide for themselves);
 * PZ_ACTIONS_SYNTH: uniform actions from the counter-based stream
 *   synth(action_seed, first_env + i, frame0 + k, agent)   (DESIGN.md "synthetic actions").
 * obs_dev (optional): observation after the last frame. stats_dev (optional) as above. */
int pz_rollout(int32_t *state_dev, int64_t n, const pz_config *cfg, int32_t K, int32_t action_source,
               uint64_t action_seed, uint64_t first_env, uint64_t frame0, void *obs_dev,
               int64_t *stats_dev, void *stream);

/* Memoised trajectory simulations for the computer players: two per-device lookup
 * tables in HBM (pz_tables_bytes() bytes, allocated by the library with cudaMalloc) holding the result
 * of calculate_expected_landing_point_x_for (physics.py:643-686) and
 * expected_landing_point_x_when_power_hit (physics.py:820-884) for every ball state with
 * |y_velocity| <= PZ_TABLE_MAX_YV, built on the device by the same simulation code. Build them explicitly
 * with pz_tables_prepare (allocates pz_tables_bytes() = 1.9 GB on the current device, runs ~0.3 s of kernels
 * and synchronises `stream`; returns 0 or the cudaError_t that prevented it — pikazoo_b200.PikaVecEnv does this
 * in its constructor). A pz_step / pz_rollout with a computer player and without PZ_FLAG_NO_TABLES that finds
 * them missing builds them implicitly the same way (one lock per device), except under stream capture, where
 * — like after a failed build — the kernels run the iterative simulations instead (results identical). */
#define PZ_TABLE_MAX_YV 100
int pz_tables_prepare(void *stream);
size_t pz_tables_bytes(void);
int pz_tables_ready(void); /* 1 if the current device holds built tables */
void pz_tables_release(void);

/* Measurement aid (bench.py `roofline.write_probe`): writes `bytes` (a multiple of 16) of device memory and nothing
 * else — the bandwidth ceiling of a write-dominated kernel, which the read+write copy figure of MEASURED_PEAKS.json
 * is not. mode 0: 128-bit stores, default cache policy; 1: with the evict-first L2 policy of the step kernel's
 * outputs; 2: st.global.cs; 3: the step kernel's observation path itself (8,960-byte blocks staged in shared memory
 * and written by cp.async.bulk with the evict-first policy, 128-thread CTAs with 35 KB of staging); 4: as 3 without
 * the policy. */
int pz_probe_write(void *dst_dev, size_t bytes, int32_t mode, void *stream);

/* packed <-> unpacked (int32 [n][53]) conversions, for checkpoints, tests and debugging */
int pz_export_state(const int32_t *state_dev, int64_t n, int32_t *unpacked_dev, void *stream);
int pz_import_state(int32_t *state_dev, int64_t n, const int32_t *unpacked_dev, void *stream);

/* ---- host-buffer path (what a numpy caller of the reference would bind) ---------------
 * A context owns the device state, staging buffers, streams and events for n envs on the
 * current device; pz_host_step copies actions host->device, runs pz_step and copies
 * obs/reward/done device->host, chunked so that copies overlap the kernel. Host buffers
 * should be pinned (cudaHostAlloc / torch pin_memory) for full PCIe speed; pageable works. */
/* chunks: env ranges stepped on their own streams (copies of one overlap the kernel of the others);
 * <= 0 picks it from n (1 for small batches, 8 from a million envs up). */
typedef struct pz_host_ctx pz_host_ctx;
int pz_host_create(pz_host_ctx **out, int64_t n, const pz_config *cfg, uint64_t base_seed,
                   uint64_t first_env, int32_t chunks);
int pz_host_reset(pz_host_ctx *ctx, void *obs_host);
int pz_host_step(pz_host_ctx *ctx, const void *actions_host, void *obs_host, void *reward_host,
                 uint8_t *done_host);
/* The same in two halves: _begin enqueues the copies and launches of every chunk and returns at once (the caller's
 * thread is free while PCIe and the GPU work), _end waits for them; the host buffers must stay valid and untouched
 * in between, and no other pz_host_* call on the context is allowed there (PZ_E_BADARG).
 * status_host (optional, uint8 [n]): pz_episode_io.status_dev — reward, terminated and truncated of an unshaped env
 * in ONE byte. With cfg->obs_layout = PZ_LAYOUT_ENV_MAJOR_SHARED, obs_dtype = PZ_OBS_I16, action_dtype = PZ_ACT_U8,
 * reward_host = done_host = NULL a step moves 73 B per env over the link instead of 305 B, with nothing lost. */
int pz_host_step_begin(pz_host_ctx *ctx, const void *actions_host, void *obs_host, void *reward_host,
                       uint8_t *done_host, uint8_t *status_host);
int pz_host_step_end(pz_host_ctx *ctx);
/* Compact WIRE format, transparent to the caller: the context keeps delivering the arrays its config names (the
 * reference's int32 — or int16 — obs [n][2][35], reward [n][2], done [n]), but what crosses the link per env is
 * player_1's int16 row + the status byte (71 B instead of 289 B); `threads` host threads (<= 0: one per hardware
 * thread) rebuild the caller's arrays chunk by chunk while the following chunks are still on the link (player_2's row
 * is a block permutation of player_1's, pikazoo_env.py:585-586; reward / done follow from the status byte when no
 * reward wrapper is fused — shaped rewards travel as they are). Needs an integer env-major observation config
 * (PZ_E_BADCONFIG otherwise). Call between steps; PZ_WIRE_NATIVE switches back. Results are identical in both modes. */
enum { PZ_WIRE_NATIVE = 0, PZ_WIRE_COMPACT = 1 };
int pz_host_set_wire(pz_host_ctx *ctx, int32_t mode, int32_t threads);
/* The host-side expansion by itself (pure host code, no device needed), for callers that receive the compact format
 * themselves (cfg->obs_layout = PZ_LAYOUT_ENV_MAJOR_SHARED + status bytes): rows int16 [n][35], status uint8 [n] ->
 * obs <int32|int16> [n][2][35], reward <f32|f64> [n][2], done uint8 [n]; any output may be NULL. */
int pz_wire_expand(const int16_t *rows, const uint8_t *status, int64_t n, int32_t obs_dtype, void *obs,
                   int32_t reward_dtype, void *reward, uint8_t *done);
size_t pz_obs_elem_bytes(int32_t obs_dtype); /* 0 for an unknown code */
int pz_host_stats(pz_host_ctx *ctx, int64_t stats_host[PZ_NUM_STATS]);
int32_t *pz_host_state_dev(pz_host_ctx *ctx);
void pz_host_destroy(pz_host_ctx *ctx);

/* ---- caller side: the MLP policy of BASELINE.json configs[4], evaluated and sampled in one kernel ----
 * The reference has no policy; this is the product's own helper for the loop `obs -> policy -> step`
 * (pikazoo_b200/policy.py), fed by the feature-major bf16 observations of pz_step (PZ_LAYOUT_FEATURE_MAJOR).
 *   obs_dev      bf16 [2][rows][ld]: element (agent, feature k, env) at (agent * rows + k) * ld + env; rows
 *                0 .. features-1 enter the contraction (policy.py folds the biases in through a row of ones)
 *   w1_dev       bf16 [2][hidden_rows][features], w2_dev bf16 [2][n_actions][w2_cols] (row-major, per agent)
 *   logits       = W2 . relu(W1 . x), fp32 accumulation, hidden activations rounded to bf16
 *   actions_dev  [n][2] of action_dtype (PZ_ACT_*): a categorical sample from softmax(logits), reproducible from
 *                the counters (seed, step, first_env + env, agent[, action]) that key its uniforms
 *                (csrc/pz_policy.cuh, restated in policy.py); greedy != 0: plain argmax
 *   logits_dev   optional fp32 [n][2][n_actions]
 *
 * Two implementations of the same network — logits bit-identical, same greedy actions: PZ_POLICY_IMPL_TCGEN05
 * (csrc/pz_policy_tc.cu: tcgen05.mma with the accumulators and the hidden activations in TMEM, tiles by TMA, one
 * thread per env in the epilogues; samples by inversion of the cumulative distribution with one uniform per env and
 * agent, policy.py inverse_cdf_reference) and PZ_POLICY_IMPL_MMA_SYNC (csrc/pz_policy.cu: warp-level mma.sync, kept
 * for A/B measurements; samples by argmax(logits + Gumbel noise), policy.py gumbel_noise_reference). pz_policy_select
 * switches process-wide and returns the previous choice, or -1 for an unknown code. */
#define PZ_POLICY_IMPL_TCGEN05 0
#define PZ_POLICY_IMPL_MMA_SYNC 1
#define PZ_POLICY_IMPL_DEFAULT PZ_POLICY_IMPL_TCGEN05
int pz_policy_select(int32_t impl);
#define PZ_POLICY_MAX_FEATURES 48
#define PZ_POLICY_MAX_HIDDEN 80
#define PZ_POLICY_MAX_ACTIONS 24
int pz_policy_mlp_act(const void *obs_dev, int64_t n, int64_t ld, int32_t rows, const void *w1_dev,
                      int32_t hidden_rows, int32_t features, const void *w2_dev, int32_t n_actions,
                      int32_t w2_cols, uint64_t seed, uint64_t step, uint64_t first_env, void *actions_dev,
                      int32_t action_dtype, int32_t greedy, float *logits_dev, void *stream);

/* ---- rgb_array rendering of selected envs ----
 * raw_env.render() / draw() (pikazoo/env/pikazoo_env.py:250-384) for n_frames frames at once: the pixel work. The host
 * (pikazoo_b200/render.py) derives, from consecutive simulation states, each frame's display list in the reference's draw
 * order — pinned against the reference's own draw() — and this rasterises them, one thread per pixel.
 *   atlas_dev       uint8 [texels][4] RGBA, every sprite variant (flipped / scaled on the host) back to back; alpha must
 *                   be 0 or 255 (true of all of the reference's sprites): compositing is a select
 *   sprites_dev     int32 [n_sprites][4] = (offset in texels, width, height, 0)
 *   background_dev  uint8 [304][432][3], the static part of the scene (draw_background, :308-339) composited once
 *   items_dev       int32 [n_frames][max_items][4] = (sprite variant, x, y, 0) back to front; variant < 0 = unused
 *   out_dev         uint8 [n_frames][304][432][3], what render() returns for render_mode "rgb_array" */
int pz_render(const uint8_t *atlas_dev, const int32_t *sprites_dev, int32_t n_sprites, const uint8_t *background_dev,
              const int32_t *items_dev, int32_t n_frames, int32_t max_items, uint8_t *out_dev, void *stream);

/* K frames of `observation -> MLP policy -> sampled actions -> raw_env.step` in ONE launch (csrc/pz_rollout_policy.cu):
 * pz_rollout with both players' actions sampled on the device from the policy above, evaluated on tcgen05 from
 * observation tiles that never leave the SM. HBM sees the packed state once per K frames.
 *   per frame k = 0 .. K-1 and env i:  o = observations of the state (bf16; NormalizeObservation applied iff
 *   cfg->normalize_observation, bit-identical to pz_step's PZ_OBS_BF16 rows; policy input row 35 = 1.0 — the bias input
 *   of policy.py's MLPPolicy —, rows 36.. = 0); logits as pz_policy_mlp_act (player_2's W1 is applied to its own
 *   permuted view of the observation; its accumulation order differs from pz_policy_mlp_act's, so logits agree to
 *   rounding, not bit for bit); action = pz_policy_mlp_act's tcgen05 sample with counters (seed, step0 + k,
 *   first_env + i, agent) or the arg-max if greedy; then one pz_rollout call of the env (NEXT-STEP auto-reset: the
 *   action sampled for a terminated env is ignored, the env is reset).
 * cfg: no computer players (PZ_E_BADCONFIG otherwise); n_actions must be the env's action space (13 with
 * cfg->simplify_action, else 18); hidden_rows <= 80, 36 <= features <= 48, w2_cols <= 80.
 *   actions_out_dev  optional uint8 [K][n][2]: the sampled actions (a trajectory buffer; the parity tests replay them
 *                    on the oracle)
 *   logits_out_dev   optional fp32 [K][n][2][n_actions] (tests)
 *   obs_dev          optional: the observation after the last frame (cfg->obs_dtype / obs_layout, as pz_observe)
 *   stats_dev        optional int64 [PZ_NUM_STATS] */
int pz_rollout_policy(int32_t *state_dev, int64_t n, const pz_config *cfg, int32_t K, const void *w1_dev,
                      int32_t hidden_rows, int32_t features, const void *w2_dev, int32_t n_actions, int32_t w2_cols,
                      uint64_t seed, uint64_t step0, uint64_t first_env, int32_t greedy, uint8_t *actions_out_dev,
                      float *logits_out_dev, void *obs_dev, int64_t *stats_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif
