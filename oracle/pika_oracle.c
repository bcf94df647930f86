/* TEST INFRASTRUCTURE — see pika_oracle.h. Plain C restatement of the reference hot path.
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Written to mirror the Python statement by statement; no attempt at speed beyond -O2.
 */
#include "pika_oracle.h"

#include <stdlib.h>
#include <string.h>

/* physics.py:9-33 */
#define GROUND_WIDTH 432
#define GROUND_HALF_WIDTH 216
#define PLAYER_LENGTH 64
#define PLAYER_HALF_LENGTH 32
#define PLAYER_TOUCHING_GROUND_Y_COORD 244
#define BALL_RADIUS 20
#define BALL_TOUCHING_GROUND_Y_COORD 252
#define NET_PILLAR_HALF_WIDTH 25
#define NET_PILLAR_TOP_TOP_Y_COORD 176
#define NET_PILLAR_TOP_BOTTOM_Y_COORD 192
#define INFINITE_LOOP_LIMIT 1000

typedef unsigned __int128 u128;

int pk_env_words(void) { return (int)(sizeof(pk_env) / 4); }

static inline int32_t iabs(int32_t v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------
 * numpy RNG (third-party; algorithm per numpy/random/bit_generator.pyx SeedSequence,
 * numpy/random/src/pcg64/pcg64.h, numpy/random/src/distributions/distributions.c
 * buffered_bounded_lemire_uint32; constants restated in SURVEY.md §8(c))
 * ---------------------------------------------------------------------------------- */
#define SS_INIT_A 0x43b0d7e5u
#define SS_MULT_A 0x931e8875u
#define SS_INIT_B 0x8b51f9ddu
#define SS_MULT_B 0x58f38dedu
#define SS_MIX_L 0xca01f9ddu
#define SS_MIX_R 0x4973f715u
#define SS_XSHIFT 16

static uint32_t ss_hashmix(uint32_t value, uint32_t *hash_const) {
    value ^= *hash_const;
    *hash_const *= SS_MULT_A;
    value *= *hash_const;
    value ^= value >> SS_XSHIFT;
    return value;
}

static uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = SS_MIX_L * x - SS_MIX_R * y;
    r ^= r >> SS_XSHIFT;
    return r;
}

/* SeedSequence(seed).generate_state(4, uint64) for a non-negative Python int seed < 2**64 */
static void ss_generate_u64x4(uint64_t seed, uint64_t out[4]) {
    uint32_t entropy[2];
    int n_entropy = 1;
    entropy[0] = (uint32_t)seed;
    entropy[1] = (uint32_t)(seed >> 32);
    if (entropy[1] != 0) n_entropy = 2; /* minimal little-endian uint32 split; 0 -> [0] */

    uint32_t pool[4];
    uint32_t hash_const = SS_INIT_A;
    for (int i = 0; i < 4; i++)
        pool[i] = ss_hashmix(i < n_entropy ? entropy[i] : 0u, &hash_const);
    for (int i_src = 0; i_src < 4; i_src++)
        for (int i_dst = 0; i_dst < 4; i_dst++)
            if (i_src != i_dst) pool[i_dst] = ss_mix(pool[i_dst], ss_hashmix(pool[i_src], &hash_const));
    /* entropy longer than the pool never happens for seeds < 2**64 */

    uint32_t words[8];
    hash_const = SS_INIT_B;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i & 3];
        v ^= hash_const;
        hash_const *= SS_MULT_B;
        v *= hash_const;
        v ^= v >> SS_XSHIFT;
        words[i] = v;
    }
    for (int i = 0; i < 4; i++) out[i] = (uint64_t)words[2 * i] | ((uint64_t)words[2 * i + 1] << 32);
}

#define PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define PCG_MULT_LO 0x4385DF649FCCF645ULL

static u128 get128(const uint32_t w[4]) {
    return ((u128)w[3] << 96) | ((u128)w[2] << 64) | ((u128)w[1] << 32) | (u128)w[0];
}
static void put128(uint32_t w[4], u128 v) {
    for (int k = 0; k < 4; k++) w[k] = (uint32_t)(v >> (32 * k));
}

void pk_pcg64_seed(uint64_t seed, uint32_t state_w[4], uint32_t inc_w[4]) {
    uint64_t s[4];
    ss_generate_u64x4(seed, s);
    const u128 mult = ((u128)PCG_MULT_HI << 64) | PCG_MULT_LO;
    u128 initstate = ((u128)s[0] << 64) | s[1];
    u128 initseq = ((u128)s[2] << 64) | s[3];
    u128 inc = (initseq << 1) | 1; /* pcg_setseq_128_srandom_r */
    u128 state = 0;
    state = state * mult + inc;
    state += initstate;
    state = state * mult + inc;
    put128(state_w, state);
    put128(inc_w, inc);
}

static uint64_t pcg64_next64(pk_env *e) {
    const u128 mult = ((u128)PCG_MULT_HI << 64) | PCG_MULT_LO;
    u128 state = get128(e->rng_state) * mult + get128(e->rng_inc);
    put128(e->rng_state, state);
    uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63)); /* XSL-RR */
}

/* pcg64_next32: low half first, high half buffered (numpy/random/src/pcg64/pcg64.h) */
static uint32_t pcg64_next32(pk_env *e) {
    if (e->has_uint32) {
        e->has_uint32 = 0;
        return e->uinteger;
    }
    uint64_t n = pcg64_next64(e);
    e->has_uint32 = 1;
    e->uinteger = (uint32_t)(n >> 32);
    return (uint32_t)n;
}

/* Generator.integers(0, high), scalar, default dtype: random_bounded_uint64 -> rng <= 2**32-1
 * -> buffered_bounded_lemire_uint32 with no local buffer. */
int32_t pk_integers(pk_env *e, uint32_t high) {
    uint32_t rng = high - 1;
    if (rng == 0) return 0;
    uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)pcg64_next32(e) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)pcg64_next32(e) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (int32_t)(m >> 32);
}

/* ------------------------------------------------------------------------------------
 * data model
 * ---------------------------------------------------------------------------------- */

/* Player.initialize_for_new_round, physics.py:181-218 */
static void player_initialize_for_new_round(pk_env *e, int i) {
    pk_player *p = &e->p[i];
    p->x = 36;
    if (i == 1) p->x = GROUND_WIDTH - 36;
    p->y = PLAYER_TOUCHING_GROUND_Y_COORD;
    p->y_velocity = 0;
    p->is_collision_with_ball_happened = 0;
    p->state = 0;
    p->frame_number = 0;
    p->normal_status_arm_swing_direction = 1;
    p->delay_before_next_frame = 0;
    p->computer_boldness = pk_integers(e, 5); /* :218 — drawn for every player, computer or not */
}

/* Ball.initialize_for_new_round, physics.py:258-277 */
static void ball_initialize_for_new_round(pk_ball *b, int is_player2_serve) {
    b->x = 56;
    if (is_player2_serve) b->x = GROUND_WIDTH - 56;
    b->y = 0;
    b->x_velocity = 0;
    b->y_velocity = 1;
    b->is_power_hit = 0;
}

/* raw_env.get_server, pikazoo_env.py:242-248 */
static int get_server(pk_env *e, const pk_config *c) {
    if (c->serve == PK_SERVE_WINNER) return e->is_player2_serve;
    if (c->serve == PK_SERVE_RANDOM) return pk_integers(e, 2) == 0;
    return (e->scores[0] + e->scores[1]) % 2 == 1;
}

/* raw_env.__init__ (pikazoo_env.py:79-111), PikaPhysics/Player/Ball/PikaUserInput
 * constructors (physics.py:43-57,143-171,224-249) followed by the S0 generator overwrite.
 * The two constructor boldness draws come from the unseeded generator and are overwritten
 * by reset(), so they are set to 0 here. */
void pk_init(pk_env *e, uint64_t seed) {
    memset(e, 0, sizeof(*e));
    for (int i = 0; i < 2; i++) {
        pk_player *p = &e->p[i];
        p->x = i ? GROUND_WIDTH - 36 : 36;
        p->y = PLAYER_TOUCHING_GROUND_Y_COORD;
        p->normal_status_arm_swing_direction = 1;
        p->diving_direction = 0;
        p->lying_down_duration_left = -1;
        p->computer_where_to_stand_by = 0;
    }
    ball_initialize_for_new_round(&e->b, 0);
    pk_pcg64_seed(seed, e->rng_state, e->rng_inc);
}

/* raw_env._get_obs/_get_player_info/_get_ball_obs, pikazoo_env.py:576-624 */
static void player_block(const pk_player *p, int32_t *o) {
    o[0] = p->x;
    o[1] = p->y;
    o[2] = p->y_velocity;
    o[3] = p->diving_direction;
    o[4] = p->lying_down_duration_left;
    o[5] = p->frame_number;
    o[6] = p->delay_before_next_frame;
    for (int s = 0; s < 5; s++) o[7 + s] = (p->state == s);
    o[12] = p->power_hit_key_is_down_previous;
}

static void get_obs(const pk_env *e, int32_t obs[70]) {
    int32_t p1[13], p2[13], ball[9];
    player_block(&e->p[0], p1);
    player_block(&e->p[1], p2);
    const pk_ball *b = &e->b;
    ball[0] = b->x;
    ball[1] = b->y;
    ball[2] = b->previous_x;
    ball[3] = b->previous_y;
    ball[4] = b->previous_previous_x;
    ball[5] = b->previous_previous_y;
    ball[6] = b->x_velocity;
    ball[7] = b->y_velocity;
    ball[8] = b->is_power_hit;
    memcpy(obs, p1, sizeof p1);
    memcpy(obs + 13, p2, sizeof p2);
    memcpy(obs + 26, ball, sizeof ball);
    memcpy(obs + 35, p2, sizeof p2);
    memcpy(obs + 48, p1, sizeof p1);
    memcpy(obs + 61, ball, sizeof ball);
}

/* raw_env.reset, pikazoo_env.py:149-173 (seed/options ignored, as in the reference) */
void pk_reset(pk_env *e, const pk_config *c, int32_t obs[70]) {
    e->game_ended = 0;
    e->round_ended = 0;
    e->is_player2_serve = 0;
    e->scores[0] = 0;
    e->scores[1] = 0;
    player_initialize_for_new_round(e, 0);
    player_initialize_for_new_round(e, 1);
    ball_initialize_for_new_round(&e->b, get_server(e, c));
    e->episode_frames = 0;
    if (obs) get_obs(e, obs);
}

/* ------------------------------------------------------------------------------------
 * physics
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t x_direction, y_direction, power_hit;
} pk_input;

/* action_key_map, pikazoo_env.py:119-141: [left, right, up, down, power_hit] */
static const uint8_t ACTION_KEY_MAP[18][5] = {
    {0, 0, 0, 0, 0}, {0, 0, 0, 0, 1}, {0, 0, 1, 0, 0}, {0, 1, 0, 0, 0}, {1, 0, 0, 0, 0},
    {0, 0, 0, 1, 0}, {0, 1, 1, 0, 0}, {1, 0, 1, 0, 0}, {0, 1, 0, 1, 0}, {1, 0, 0, 1, 0},
    {0, 0, 1, 0, 1}, {0, 1, 0, 0, 1}, {1, 0, 0, 0, 1}, {0, 0, 0, 1, 1}, {0, 1, 1, 0, 1},
    {1, 0, 1, 0, 1}, {0, 1, 0, 1, 1}, {1, 0, 0, 1, 1},
};

/* PikaUserInput.get_input, physics.py:59-99 (rows are 5 wide => down_right_key is None) */
static void get_input(pk_player *p, const uint8_t key[5], pk_input *in) {
    int left = key[0], right = key[1], up = key[2], down = key[3], power = key[4];
    if (left)
        in->x_direction = -1;
    else if (right)
        in->x_direction = 1;
    else
        in->x_direction = 0;
    if (up)
        in->y_direction = -1;
    else if (down)
        in->y_direction = 1;
    else
        in->y_direction = 0;
    in->power_hit = (!p->power_hit_key_is_down_previous && power) ? 1 : 0;
    p->power_hit_key_is_down_previous = power;
}

/* process_collision_between_ball_and_world_and_set_ball_position, physics.py:359-436 */
static int ball_world(pk_ball *b) {
    b->previous_previous_x = b->previous_x;
    b->previous_previous_y = b->previous_y;
    b->previous_x = b->x;
    b->previous_y = b->y;
    /* :373-388 fine_rotation/rotation are render-only */
    int32_t future_ball_x = b->x + b->x_velocity;
    if (future_ball_x < BALL_RADIUS || future_ball_x > GROUND_WIDTH) b->x_velocity = -b->x_velocity;
    int32_t future_ball_y = b->y + b->y_velocity;
    if (future_ball_y < 0) b->y_velocity = 1;
    if (iabs(b->x - GROUND_HALF_WIDTH) < NET_PILLAR_HALF_WIDTH && b->y > NET_PILLAR_TOP_TOP_Y_COORD) {
        if (b->y <= NET_PILLAR_TOP_BOTTOM_Y_COORD) {
            if (b->y_velocity > 0) b->y_velocity = -b->y_velocity;
        } else {
            if (b->x < GROUND_HALF_WIDTH)
                b->x_velocity = -iabs(b->x_velocity);
            else
                b->x_velocity = iabs(b->x_velocity);
        }
    }
    future_ball_y = b->y + b->y_velocity;
    if (future_ball_y > BALL_TOUCHING_GROUND_Y_COORD) {
        b->y_velocity = -b->y_velocity;
        b->punch_effect_x = b->x;
        b->y = BALL_TOUCHING_GROUND_Y_COORD;
        return 1;
    }
    b->y = future_ball_y;
    b->x = b->x + b->x_velocity;
    b->y_velocity += 1;
    return 0;
}

/* calculate_expected_landing_point_x_for, physics.py:643-686 */
static void calculate_expected_landing_point_x_for(pk_ball *b) {
    int32_t x = b->x, y = b->y, xv = b->x_velocity, yv = b->y_velocity;
    int loop_counter = 0;
    for (;;) {
        loop_counter += 1;
        int32_t future_x = xv + x;
        if (future_x < BALL_RADIUS || future_x > GROUND_WIDTH) xv = -xv;
        if (y + yv < 0) yv = 1;
        if (iabs(x - GROUND_HALF_WIDTH) < NET_PILLAR_HALF_WIDTH && y > NET_PILLAR_TOP_TOP_Y_COORD) {
            if (y < NET_PILLAR_TOP_BOTTOM_Y_COORD) { /* strict, unlike ball_world */
                if (yv > 0) yv = -yv;
            } else {
                if (x < GROUND_HALF_WIDTH)
                    xv = -iabs(xv);
                else
                    xv = iabs(xv);
            }
        }
        y = y + yv;
        if (y > BALL_TOUCHING_GROUND_Y_COORD || loop_counter >= INFINITE_LOOP_LIMIT) break;
        x = x + xv;
        yv += 1;
    }
    b->expected_landing_point_x = x;
}

/* the loop of expected_landing_point_x_when_power_hit, physics.py:847-884 */
static int32_t power_hit_loop(int32_t x, int32_t y, int32_t xv, int32_t yv) {
    int loop_counter = 0;
    for (;;) {
        loop_counter += 1;
        int32_t future_x = x + xv;
        if (future_x < BALL_RADIUS || future_x > GROUND_WIDTH) xv = -xv;
        if (y + yv < 0) yv = 1;
        if (iabs(x - GROUND_HALF_WIDTH) < NET_PILLAR_HALF_WIDTH && y > NET_PILLAR_TOP_TOP_Y_COORD) {
            if (yv > 0) yv = -yv; /* :865-866 whole net zone just bounces */
        }
        y = y + yv;
        if (y > BALL_TOUCHING_GROUND_Y_COORD || loop_counter >= INFINITE_LOOP_LIMIT) return x;
        x = x + xv;
        yv += 1;
    }
}

/* expected_landing_point_x_when_power_hit, physics.py:820-884 */
static int32_t expected_landing_point_x_when_power_hit(int x_dir, int y_dir, const pk_ball *b) {
    int32_t xv;
    if (b->x < GROUND_HALF_WIDTH)
        xv = (iabs(x_dir) + 1) * 10;
    else
        xv = -(iabs(x_dir) + 1) * 10;
    return power_hit_loop(b->x, b->y, xv, iabs(b->y_velocity) * y_dir * 2);
}

/* test helper: the two trajectory loops on n arbitrary (x, y, x_velocity, y_velocity) starts */
void pk_simulate_many(int64_t n, const int32_t *xyv, int power, int32_t *out) {
    for (int64_t i = 0; i < n; i++) {
        const int32_t *q = xyv + 4 * i;
        if (power) {
            out[i] = power_hit_loop(q[0], q[1], q[2], q[3]);
        } else {
            pk_ball b;
            b.x = q[0], b.y = q[1], b.x_velocity = q[2], b.y_velocity = q[3];
            calculate_expected_landing_point_x_for(&b);
            out[i] = b.expected_landing_point_x;
        }
    }
}

/* decide_whether_input_power_hit, physics.py:774-817 */
static int decide_whether_input_power_hit(pk_env *e, int i, pk_input *in) {
    const pk_player *p = &e->p[i];
    const pk_player *other = &e->p[1 - i];
    int is_p2 = i;
    int first = pk_integers(e, 2) == 0; /* :795 */
    for (int x_direction = 1; x_direction > -1; x_direction--) {
        for (int k = 0; k < 3; k++) {
            int y_direction = first ? (-1 + k) : (1 - k); /* range(-1,2,1) vs range(1,-2,-1) */
            int32_t lx = expected_landing_point_x_when_power_hit(x_direction, y_direction, &e->b);
            if ((lx <= is_p2 * GROUND_HALF_WIDTH || lx >= is_p2 * GROUND_WIDTH + GROUND_HALF_WIDTH) &&
                iabs(lx - other->x) > PLAYER_LENGTH) {
                in->x_direction = x_direction;
                in->y_direction = y_direction;
                return 1;
            }
        }
    }
    (void)p;
    return 0;
}

/* let_computer_decide_user_input, physics.py:689-771 */
static void let_computer_decide_user_input(pk_env *e, int i, pk_input *in) {
    pk_player *p = &e->p[i];
    const pk_player *other = &e->p[1 - i];
    const pk_ball *b = &e->b;
    int is_p2 = i;
    in->x_direction = 0;
    in->y_direction = 0;
    in->power_hit = 0;

    int32_t virtual_lx = b->expected_landing_point_x;
    if (iabs(b->x - p->x) > 100 && iabs(b->x_velocity) < p->computer_boldness + 5) {
        int32_t left_boundary = is_p2 * GROUND_HALF_WIDTH;
        if ((b->expected_landing_point_x <= left_boundary ||
             b->expected_landing_point_x >= is_p2 * GROUND_WIDTH + GROUND_HALF_WIDTH) &&
            p->computer_where_to_stand_by == 0) {
            virtual_lx = left_boundary + (GROUND_HALF_WIDTH / 2);
        }
    }

    if (iabs(virtual_lx - p->x) > p->computer_boldness + 8) {
        if (p->x < virtual_lx)
            in->x_direction = 1;
        else
            in->x_direction = -1;
    } else if (pk_integers(e, 20) == 0) {                 /* :728 */
        p->computer_where_to_stand_by = pk_integers(e, 2); /* :729 */
    }

    if (p->state == 0) {
        if (iabs(b->x_velocity) < p->computer_boldness + 3 && iabs(b->x - p->x) < PLAYER_HALF_LENGTH &&
            b->y > -36 && b->y < 10 * p->computer_boldness + 84 && b->y_velocity > 0) {
            in->y_direction = -1;
        }
        int32_t left_boundary = is_p2 * GROUND_HALF_WIDTH;
        int32_t right_boundary = (is_p2 + 1) * GROUND_HALF_WIDTH;
        if (b->expected_landing_point_x > left_boundary && b->expected_landing_point_x < right_boundary &&
            iabs(b->x - p->x) > p->computer_boldness * 5 + PLAYER_LENGTH && b->x > left_boundary &&
            b->x < right_boundary && b->y > 174) {
            in->power_hit = 1;
            if (p->x < b->x)
                in->x_direction = 1;
            else
                in->x_direction = -1;
        }
    } else if (p->state == 1 || p->state == 2) {
        if (iabs(b->x - p->x) > 8) {
            if (p->x < b->x)
                in->x_direction = 1;
            else
                in->x_direction = -1;
        }
        if (iabs(b->x - p->x) < 48 && iabs(b->y - p->y) < 48) {
            int will = decide_whether_input_power_hit(e, i, in);
            if (will) {
                in->power_hit = 1;
                if (iabs(other->x - p->x) < 80 && in->y_direction != -1) in->y_direction = -1;
            }
        }
    }
}

/* process_player_movement_and_set_player_position, physics.py:439-564 */
static void player_movement(pk_env *e, const pk_config *c, int i, pk_input *in) {
    pk_player *p = &e->p[i];
    int is_computer = i ? c->is_player2_computer : c->is_player1_computer;
    if (is_computer) let_computer_decide_user_input(e, i, in);

    if (p->state == 4) {
        p->lying_down_duration_left += -1;
        if (p->lying_down_duration_left < -1) p->state = 0;
        return;
    }

    int32_t player_velocity_x = 0;
    if (p->state < 5) {
        if (p->state < 3)
            player_velocity_x = in->x_direction * 6;
        else
            player_velocity_x = p->diving_direction * 8;
    }
    int32_t future_player_x = p->x + player_velocity_x;
    p->x = future_player_x;
    if (i == 0) {
        if (future_player_x < PLAYER_HALF_LENGTH)
            p->x = PLAYER_HALF_LENGTH;
        else if (future_player_x > GROUND_HALF_WIDTH - PLAYER_HALF_LENGTH)
            p->x = GROUND_HALF_WIDTH - PLAYER_HALF_LENGTH;
    } else {
        if (future_player_x < GROUND_HALF_WIDTH + PLAYER_HALF_LENGTH)
            p->x = GROUND_HALF_WIDTH + PLAYER_HALF_LENGTH;
        else if (future_player_x > GROUND_WIDTH - PLAYER_HALF_LENGTH)
            p->x = GROUND_WIDTH - PLAYER_HALF_LENGTH;
    }

    if (p->state < 3 && in->y_direction == -1 && p->y == PLAYER_TOUCHING_GROUND_Y_COORD) {
        p->y_velocity = -16;
        p->state = 1;
        p->frame_number = 0;
    }

    int32_t future_player_y = p->y + p->y_velocity;
    p->y = future_player_y;
    if (future_player_y < PLAYER_TOUCHING_GROUND_Y_COORD) {
        p->y_velocity += 1;
    } else if (future_player_y > PLAYER_TOUCHING_GROUND_Y_COORD) {
        p->y_velocity = 0;
        p->y = PLAYER_TOUCHING_GROUND_Y_COORD;
        p->frame_number = 0;
        if (p->state == 3) {
            p->state = 4;
            p->frame_number = 0;
            p->lying_down_duration_left = 3;
        } else {
            p->state = 0;
        }
    }

    if (in->power_hit == 1) {
        if (p->state == 1) {
            p->delay_before_next_frame = 5;
            p->frame_number = 0;
            p->state = 2;
        } else if (p->state == 0 && in->x_direction != 0) {
            p->state = 3;
            p->frame_number = 0;
            p->diving_direction = in->x_direction;
            p->y_velocity = -5;
        }
    }

    if (p->state == 1) {
        p->frame_number = (p->frame_number + 1) % 3;
    } else if (p->state == 2) {
        if (p->delay_before_next_frame < 1) {
            p->frame_number += 1;
            if (p->frame_number > 4) {
                p->frame_number = 0;
                p->state = 1;
            }
        } else {
            p->delay_before_next_frame -= 1;
        }
    } else if (p->state == 0) {
        p->delay_before_next_frame += 1;
        if (p->delay_before_next_frame > 3) {
            p->delay_before_next_frame = 0;
            int32_t future_frame_number = p->frame_number + p->normal_status_arm_swing_direction;
            if (future_frame_number < 0 || future_frame_number > 4)
                p->normal_status_arm_swing_direction = -p->normal_status_arm_swing_direction;
            p->frame_number = p->frame_number + p->normal_status_arm_swing_direction;
        }
    }
    /* :554-564 player.game_ended block: unreachable before termination */
}

/* is_collision_between_ball_and_player_happened, physics.py:340-356 */
static int is_collision(const pk_ball *b, int32_t px, int32_t py) {
    if (iabs(b->x - px) <= PLAYER_HALF_LENGTH)
        if (iabs(b->y - py) <= PLAYER_HALF_LENGTH) return 1;
    return 0;
}

/* process_collision_between_ball_and_player, physics.py:580-640 */
static void collide_ball_player(pk_env *e, int32_t player_x, const pk_input *in, int32_t player_state) {
    pk_ball *b = &e->b;
    if (b->x < player_x)
        b->x_velocity = -(iabs(b->x - player_x) / 3);
    else if (b->x > player_x)
        b->x_velocity = iabs(b->x - player_x) / 3;
    if (b->x_velocity == 0) b->x_velocity = pk_integers(e, 3) - 1; /* :613 */
    int32_t ball_abs_y_velocity = iabs(b->y_velocity);
    b->y_velocity = -ball_abs_y_velocity;
    if (ball_abs_y_velocity < 15) b->y_velocity = -15;
    if (player_state == 2) {
        if (b->x < GROUND_HALF_WIDTH)
            b->x_velocity = (iabs(in->x_direction) + 1) * 10;
        else
            b->x_velocity = -(iabs(in->x_direction) + 1) * 10;
        b->punch_effect_x = b->x; /* :628 (punch_effect_y/radius are render-only) */
        b->y_velocity = iabs(b->y_velocity) * in->y_direction * 2;
        b->is_power_hit = 1;
    } else {
        b->is_power_hit = 0;
    }
}

/* physics_engine, physics.py:280-337 */
static int physics_engine(pk_env *e, const pk_config *c, pk_input in[2]) {
    int any_computer = c->is_player1_computer || c->is_player2_computer;
    int is_ball_touching_ground = ball_world(&e->b);
    for (int i = 0; i < 2; i++) {
        if (any_computer) calculate_expected_landing_point_x_for(&e->b);
        player_movement(e, c, i, &in[i]);
    }
    for (int i = 0; i < 2; i++) {
        pk_player *p = &e->p[i];
        if (is_collision(&e->b, p->x, p->y)) {
            if (!p->is_collision_with_ball_happened) {
                collide_ball_player(e, p->x, &in[i], p->state);
                if (any_computer) calculate_expected_landing_point_x_for(&e->b);
                p->is_collision_with_ball_happened = 1;
            }
        } else {
            p->is_collision_with_ball_happened = 0;
        }
    }
    return is_ball_touching_ground;
}

/* SimplifyAction.action_map, simplify_action.py:16-19 */
static const int8_t SIMPLIFY_MAP[2][13] = {
    {0, 1, 2, 3, 4, 6, 7, 10, 11, 12, 13, 14, 16},
    {0, 1, 2, 4, 3, 7, 6, 10, 12, 11, 13, 15, 17},
};

/* [RewardByBallPosition.step (reward_by_ball_position.py:20-31) o SimplifyAction.step
 * (simplify_action.py:22-25) o] raw_env.step (pikazoo_env.py:175-240) */
int pk_step(pk_env *e, const pk_config *c, int32_t a1, int32_t a2, int32_t obs[70], double reward[2],
            uint8_t *terminated) {
    int32_t a[2] = {a1, a2};
    for (int i = 0; i < 2; i++) {
        if (c->simplify_action) {
            if (a[i] < 0 || a[i] >= 13) return -1;
            a[i] = SIMPLIFY_MAP[i][a[i]];
        }
        if (a[i] < 0 || a[i] >= 18) return -1; /* numpy would wrap negatives; the product rejects them */
    }

    if (e->round_ended && !e->game_ended) { /* :176-180 */
        player_initialize_for_new_round(e, 0);
        player_initialize_for_new_round(e, 1);
        ball_initialize_for_new_round(&e->b, get_server(e, c));
        e->round_ended = 0;
    }

    pk_input in[2];
    for (int i = 0; i < 2; i++) get_input(&e->p[i], ACTION_KEY_MAP[a[i]], &in[i]); /* :182-184 */

    int is_ball_touching_ground = physics_engine(e, c, in); /* :186 */

    if (is_ball_touching_ground && !e->round_ended && !e->game_ended) { /* :190-210 */
        if (e->b.punch_effect_x < GROUND_HALF_WIDTH) {
            e->is_player2_serve = 1;
            e->scores[1] += 1;
            if (e->scores[1] >= c->winning_score) e->game_ended = 1;
        } else {
            e->is_player2_serve = 0;
            e->scores[0] += 1;
            if (e->scores[0] >= c->winning_score) e->game_ended = 1;
        }
        e->round_ended = 1;
    }
    e->episode_frames += 1;

    if (obs) get_obs(e, obs); /* :215 */

    int player1_reward = 0; /* :217-223 */
    if (e->round_ended) player1_reward = e->is_player2_serve ? -1 : 1;
    if (reward) {
        reward[0] = (double)player1_reward;
        reward[1] = (double)(-player1_reward);
        /* wrappers, innermost first. RewardInNormalState.step, reward_in_normal_state.py:10-15 */
        if (c->reward_in_normal_state == 2)
            for (int i = 0; i < 2; i++)
                if (reward[i] == 0) reward[i] = c->normal_state_reward;
        if (c->reward_by_ball_position) { /* RewardByBallPosition.step, reward_by_ball_position.py:20-31 */
            int x_sign = e->b.x >= c->x_line; /* obs["player_1"][26], [27] */
            int y_sign = e->b.y > c->y_line;
            int ball_pos = 1 * y_sign + 2 * x_sign;
            for (int i = 0; i < 2; i++) reward[i] += c->additional_reward[i * 4 + ball_pos];
        }
        if (c->reward_in_normal_state == 1)
            for (int i = 0; i < 2; i++)
                if (reward[i] == 0) reward[i] = c->normal_state_reward;
    }
    if (terminated) *terminated = (uint8_t)e->game_ended; /* :233 */
    return 0;
}

/* ------------------------------------------------------------------------------------
 * batched helpers (product semantics, DESIGN.md "auto-reset")
 * ---------------------------------------------------------------------------------- */
void pk_vec_init(pk_env *envs, int64_t n, uint64_t base_seed) {
    for (int64_t i = 0; i < n; i++) pk_init(&envs[i], base_seed + (uint64_t)i);
}

void pk_vec_reset(pk_env *envs, int64_t n, const pk_config *c, int32_t *obs) {
    for (int64_t i = 0; i < n; i++) pk_reset(&envs[i], c, obs ? obs + 70 * i : NULL);
}

/* raw_env._get_obs of every env (pikazoo_env.py:576-624) without stepping */
void pk_vec_obs(const pk_env *envs, int64_t n, int32_t *obs) {
    for (int64_t i = 0; i < n; i++) get_obs(&envs[i], obs + 70 * i);
}

static int is_truncated(const pk_env *e, const pk_config *c) {
    return c->max_episode_frames > 0 && !e->game_ended && e->episode_frames >= c->max_episode_frames;
}

void pk_vec_reset_ex(pk_env *envs, int64_t n, const pk_config *c, int32_t *obs, double *ep_return,
                     int32_t *ep_length) {
    pk_vec_reset(envs, n, c, obs);
    for (int64_t i = 0; i < n; i++) { /* record_episode_statistics.py:20-26 */
        if (ep_return) ep_return[2 * i] = ep_return[2 * i + 1] = 0.0;
        if (ep_length) ep_length[i] = 0;
    }
}

int pk_vec_step_ex(pk_env *envs, int64_t n, const pk_config *c, const int32_t *actions, int32_t *obs,
                   double *reward, uint8_t *done, int autoreset, double *ep_return,
                   int32_t *ep_length, uint8_t *truncated) {
    int rc = 0;
    for (int64_t i = 0; i < n; i++) {
        pk_env *e = &envs[i];
        int32_t *o = obs ? obs + 70 * i : NULL;
        if (e->game_ended || is_truncated(e, c)) {
            if (autoreset) {
                pk_reset(e, c, o);
                if (ep_return) ep_return[2 * i] = ep_return[2 * i + 1] = 0.0; /* :24-25 */
            } else {
                if (o) get_obs(e, o);
            }
            if (reward) reward[2 * i] = reward[2 * i + 1] = 0.0;
        } else {
            uint8_t t = 0;
            double r[2];
            if (pk_step(e, c, actions ? actions[2 * i] : 0, actions ? actions[2 * i + 1] : 0, o, r, &t) != 0) rc = -1;
            if (reward) {
                reward[2 * i] = r[0];
                reward[2 * i + 1] = r[1];
            }
            if (ep_return) { /* :32 */
                ep_return[2 * i] += r[0];
                ep_return[2 * i + 1] += r[1];
            }
        }
        if (done) done[i] = (uint8_t)e->game_ended;
        if (ep_length) ep_length[i] = e->episode_frames; /* :33 */
        if (truncated) truncated[i] = (uint8_t)is_truncated(e, c);
    }
    return rc;
}

int pk_vec_step(pk_env *envs, int64_t n, const pk_config *c, const int32_t *actions, int32_t *obs,
                double *reward, uint8_t *done, int autoreset) {
    return pk_vec_step_ex(envs, n, c, actions, obs, reward, done, autoreset, NULL, NULL, NULL);
}

/* raw_env.observation_space bounds, pikazoo_env.py:485-562 */
static const int32_t OBS_LOW[35] = {32, 108, -15, -1, -2, 0, 0, 0, 0, 0, 0, 0, 0,
                                    32, 108, -15, -1, -2, 0, 0, 0, 0, 0, 0, 0, 0,
                                    20, 0, 0, 0, 0, 0, -20, -124, 0};
static const int32_t OBS_HIGH[35] = {400, 244, 16, 1, 3, 4, 4, 1, 1, 1, 1, 1, 1,
                                     400, 244, 16, 1, 3, 4, 4, 1, 1, 1, 1, 1, 1,
                                     432, 252, 432, 252, 432, 252, 20, 124, 1};

/* NormalizeObservation.step / reset, normalize_observation.py:18-32: numpy true division of two
 * int64 arrays = one IEEE double division per element */
void pk_normalize_obs(int64_t n, const int32_t *obs, double *out) {
    for (int64_t i = 0; i < n; i++)
        for (int a = 0; a < 2; a++)
            for (int k = 0; k < 35; k++) {
                int64_t j = i * 70 + a * 35 + k;
                out[j] = (double)(obs[j] - OBS_LOW[k]) / (double)(OBS_HIGH[k] - OBS_LOW[k]);
            }
}

/* splitmix64-style finaliser over (seed, env, frame, agent); product-defined, DESIGN.md */
int32_t pk_synth_action(uint64_t action_seed, uint64_t global_env, uint64_t frame, int agent,
                        uint32_t n_actions) {
    uint64_t z = action_seed + 0x9E3779B97F4A7C15ULL * (2 * global_env + (uint64_t)agent + 1);
    z ^= frame * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int32_t)(((z >> 32) * (uint64_t)n_actions) >> 32);
}

int64_t pk_vec_rollout(pk_env *envs, int64_t n, const pk_config *c, int K, int action_mode,
                       uint64_t action_seed, uint64_t first_env, uint64_t frame0, int64_t *stats) {
    int64_t local[9] = {0};
    uint32_t n_actions = c->simplify_action ? 13u : 18u;
    for (int64_t i = 0; i < n; i++) {
        pk_env *e = &envs[i];
        for (int k = 0; k < K; k++) {
            if (e->game_ended || is_truncated(e, c)) {
                pk_reset(e, c, NULL);
                local[7] += 1;
                continue;
            }
            int32_t a1 = 0, a2 = 0;
            if (action_mode == 1) {
                a1 = pk_synth_action(action_seed, first_env + (uint64_t)i, frame0 + (uint64_t)k, 0, n_actions);
                a2 = pk_synth_action(action_seed, first_env + (uint64_t)i, frame0 + (uint64_t)k, 1, n_actions);
            }
            uint8_t t = 0;
            pk_step(e, c, a1, a2, NULL, NULL, &t);
            local[0] += 1;
            if (e->round_ended) { /* set on exactly the frame a point is scored (cleared at :180) */
                if (e->is_player2_serve)
                    local[6] += 1;
                else
                    local[5] += 1;
            }
            if (t) {
                local[1] += 1;
                local[2] += e->episode_frames;
                if (e->scores[0] > e->scores[1])
                    local[3] += 1;
                else
                    local[4] += 1;
            } else if (is_truncated(e, c)) {
                local[8] += 1;
            }
        }
    }
    if (stats)
        for (int k = 0; k < 9; k++) stats[k] += local[k];
    return (int64_t)n * K;
}
