// One frame of one env, entirely in registers: the device form of
//   raw_env.step / reset                 pikazoo/env/pikazoo_env.py:149-248
//   PikaUserInput.get_input              pikazoo/env/physics.py:59-99
//   physics_engine and its callees       pikazoo/env/physics.py:280-884
// One thread owns one env. The divergent trajectory-simulation loops and the computer's
// power-hit search are written as warp-collective loops over the mask of lanes that step
// this frame (warp votes keep the warp converged; a lane that has finished idles predicated
// instead of forcing a reconvergence stack).
#pragma once
#include <utility>

#include "pz_rng.cuh"

namespace pz {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// Memoised trajectory simulations (see simulate_landing_x).
//   land [yv + kTabYv][xv + 20][y][x]      = landing x | ended-on-ground << 15    yv in [-kTabYv, kTabYv], xv in [-20, 20]
//   power[x_direction][yv0 / 2 + kTabYv][y][x] = landing x                        yv0 = |ball yv| * y_direction * 2
// with y in [0, 252] and x in [0, 432].
constexpr int kTabYv = 100;
constexpr int kTabNx = kGroundWidth + 1, kTabNy = kBallGroundY + 1, kTabNxv = 41, kTabNyv = 2 * kTabYv + 1;
constexpr int64_t kTabLandEntries = (int64_t)kTabNyv * kTabNxv * kTabNy * kTabNx;
constexpr int64_t kTabPowerEntries = (int64_t)2 * kTabNyv * kTabNy * kTabNx;

__host__ __device__ __forceinline__ int tab_land_index(int x, int y, int xv, int yv) {
    return (((yv + kTabYv) * kTabNxv + (xv + 20)) * kTabNy + y) * kTabNx + x;
}
__host__ __device__ __forceinline__ int tab_power_index(int x, int y, int xd, int half_yv0) {
    return ((xd * kTabNyv + (half_yv0 + kTabYv)) * kTabNy + y) * kTabNx + x;
}
__device__ __forceinline__ bool tab_pos_ok(int x, int y) {
    return (unsigned)x < (unsigned)kTabNx && (unsigned)y < (unsigned)kTabNy;
}

// Per-launch constants (uniform over the grid)
struct StepCfg {
    int winning_score;
    int serve;  // PZ_SERVE_*
    const uint16_t *tab_land;   // nullptr: iterate
    const uint16_t *tab_power;
};

// PCG64 stream of the env owned by this thread. LAZY: loaded from global memory on the first draw
// (kernels without computer players, where most frames draw nothing); otherwise the caller has loaded it
// and neither the state pointers nor the env index stay live across the frame (10+ registers in the
// rollout kernel).
template <bool LAZY>
struct DrawCtxT {
    Rng r;
    StatePtrs s;
    int64_t idx;
    __device__ __forceinline__ void ensure() {
        if (LAZY && !r.loaded) rng_load(r, s, idx);
    }
    template <uint32_t HIGH>
    __device__ __forceinline__ int integers(int &has32) {
        ensure();
        return rng_integers<HIGH>(r, has32);
    }
    __device__ __forceinline__ void computer_draws(int &has32, bool near, bool search, int &standby, int &y_first) {
        ensure();
        rng_computer_draws(r, has32, near, search, standby, y_first);
    }
    __device__ __forceinline__ void integers5_twice(int &has32, int &a, int &b) {
        ensure();
        rng_integers5_twice(r, has32, a, b);
    }
};
using DrawCtx = DrawCtxT<true>;

// ---- action decode -------------------------------------------------------------------------
// action_key_map (pikazoo_env.py:119-141) as key-bit masks: bit0 left, bit1 right, bit2 up,
// bit3 down, bit4 power_hit. Entry 0 is all-zero, entries 1..12 / 13..17 are packed 5 bits each.
constexpr uint32_t kL = 1, kR = 2, kU = 4, kD = 8, kP = 16;
constexpr uint32_t kKeyMap[18] = {0,       kP,      kU,      kR,           kL,           kD,
                                  kR | kU, kL | kU, kR | kD, kL | kD,      kU | kP,      kR | kP,
                                  kL | kP, kD | kP, kR | kU | kP, kL | kU | kP, kR | kD | kP, kL | kD | kP};
// SimplifyAction.action_map (simplify_action.py:16-19)
constexpr int kSimplify[2][13] = {{0, 1, 2, 3, 4, 6, 7, 10, 11, 12, 13, 14, 16},
                                  {0, 1, 2, 4, 3, 7, 6, 10, 12, 11, 13, 15, 17}};

__host__ __device__ constexpr uint64_t pack_keys_18_lo() {
    uint64_t v = 0;
    for (int a = 1; a <= 12; a++) v |= (uint64_t)kKeyMap[a] << (5 * (a - 1));
    return v;
}
__host__ __device__ constexpr uint64_t pack_keys_18_hi() {
    uint64_t v = 0;
    for (int a = 13; a <= 17; a++) v |= (uint64_t)kKeyMap[a] << (5 * (a - 13));
    return v;
}
__host__ __device__ constexpr uint64_t pack_keys_13(int agent) {
    uint64_t v = 0;
    for (int a = 1; a <= 12; a++) v |= (uint64_t)kKeyMap[kSimplify[agent][a]] << (5 * (a - 1));
    return v;
}

struct Input {
    int xdir, ydir, power;
};

// Returns the 5 key bits of `action` for player I; *bad is set when the action is outside
// the action space (the reference would raise IndexError; here it is counted and treated as 0).
template <int I, bool SIMPLIFY>
__device__ __forceinline__ uint32_t decode_keys(int action, bool &bad) {
    constexpr int n_actions = SIMPLIFY ? 13 : 18;
    bad = (unsigned)action >= (unsigned)n_actions;
    if (bad || action == 0) return 0;
    if (SIMPLIFY) {
        constexpr uint64_t t = pack_keys_13(I);
        return (uint32_t)(t >> (5 * (action - 1))) & 31u;
    } else {
        constexpr uint64_t lo = pack_keys_18_lo(), hi = pack_keys_18_hi();
        return action <= 12 ? ((uint32_t)(lo >> (5 * (action - 1))) & 31u)
                            : ((uint32_t)(hi >> (5 * (action - 13))) & 31u);
    }
}


// PikaUserInput.get_input, physics.py:59-99 (rows are 5 wide: down_right_key is None)
__device__ __forceinline__ Input get_input(Player &p, uint32_t keys) {
    Input in;
    in.xdir = (keys & kL) ? -1 : ((keys & kR) ? 1 : 0);
    in.ydir = (keys & kU) ? -1 : ((keys & kD) ? 1 : 0);
    int down = (keys & kP) ? 1 : 0;
    in.power = (down && !p.keyprev) ? 1 : 0;
    p.keyprev = down;
    return in;
}

// action -> Input in one go (action_key_map + get_input): five 18-bit masks, one per key, bit a = "action a holds
// the key" — a shift and an AND per key instead of a 64-bit table shift, its select and the key-bit tests of
// get_input (the two decodes were 6-7 % of a step kernel's instructions). No action holds left AND right or up AND
// down, so the directions are differences of bits. Identical to get_input(p, decode_keys(action)) for every action,
// in and out of range (tests/test_device_code_on_host.py).
template <int I, bool SIMPLIFY>
__host__ __device__ constexpr uint32_t action_mask(uint32_t key) {
    uint32_t m = 0;
    for (int a = 0; a < (SIMPLIFY ? 13 : 18); a++)
        if (kKeyMap[SIMPLIFY ? kSimplify[I][a] : a] & key) m |= 1u << a;
    return m;
}
template <int I, bool SIMPLIFY>
__device__ __forceinline__ Input decode_input(int action, Player &p, bool &bad) {
    constexpr int n_actions = SIMPLIFY ? 13 : 18;
    constexpr uint32_t mL = action_mask<I, SIMPLIFY>(kL), mR = action_mask<I, SIMPLIFY>(kR),
                       mU = action_mask<I, SIMPLIFY>(kU), mD = action_mask<I, SIMPLIFY>(kD),
                       mP = action_mask<I, SIMPLIFY>(kP);
    bad = (unsigned)action >= (unsigned)n_actions;
    const uint32_t a = bad ? 0u : (uint32_t)action;  // out of range: counted, no key held
    Input in;
    in.xdir = (int)((mR >> a) & 1u) - (int)((mL >> a) & 1u);
    in.ydir = (int)((mD >> a) & 1u) - (int)((mU >> a) & 1u);
    const int down = (int)((mP >> a) & 1u);
    in.power = down & (p.keyprev ^ 1);
    p.keyprev = down;
    return in;
}

// ---- round / game initialisation -------------------------------------------------------------
// Player.initialize_for_new_round, physics.py:181-218 (computer_boldness, :218, is drawn by new_round)
template <int I>
__device__ __forceinline__ void player_new_round(Env &e) {
    Player &p = e.p[I];
    p.x = I ? kGroundWidth - 36 : 36;
    p.y = kPlayerGroundY;
    p.yv = 0;
    p.coll = 0;
    p.state = 0;
    p.frame = 0;
    p.arm = 1;
    p.delay = 0;
}

// Ball.initialize_for_new_round (physics.py:258-277) with raw_env.get_server (pikazoo_env.py:242-248)
template <class Ctx>
__device__ __forceinline__ void new_round(Env &e, Ctx &d, const StepCfg &c) {
    player_new_round<0>(e);
    player_new_round<1>(e);
    d.integers5_twice(e.has32, e.p[0].bold, e.p[1].bold);  // :218, player 1 then player 2
    int p2_serves;
    if (c.serve == 0)
        p2_serves = e.p2serve;
    else if (c.serve == 2)
        p2_serves = d.template integers<2>(e.has32) == 0;
    else
        p2_serves = ((e.score[0] + e.score[1]) & 1);
    Ball &b = e.b;
    b.x = p2_serves ? kGroundWidth - 56 : 56;
    b.y = 0;
    b.xv = 0;
    b.yv = 1;
    b.pow = 0;
    e.land_ok = 0;  // the ball was re-placed: the cached landing point is stale
}

// raw_env.reset, pikazoo_env.py:149-173 — on a live object: everything not assigned here
// carries over (previous positions, landing point, diving direction, ..., the RNG stream).
template <class Ctx>
__device__ __forceinline__ void reset_env(Env &e, Ctx &d, const StepCfg &c) {
    e.game_ended = 0;
    e.round_ended = 0;
    e.p2serve = 0;
    e.score[0] = 0;
    e.score[1] = 0;
    new_round(e, d, c);
    e.ep_frames = 0;
}

// Constructor defaults of a fresh reference env (physics.py:43-57,143-171,224-249;
// pikazoo_env.py:100-111). Boldness is 0 until reset() draws it.
__device__ __forceinline__ void fresh_env(Env &e) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        Player &p = e.p[i];
        p.x = i ? kGroundWidth - 36 : 36;
        p.y = kPlayerGroundY;
        p.yv = 0;
        p.state = 0;
        p.frame = 0;
        p.delay = 0;
        p.arm = 1;
        p.dive = 0;
        p.lying = -1;
        p.coll = 0;
        p.bold = 0;
        p.standby = 0;
        p.keyprev = 0;
    }
    Ball &b = e.b;
    b.x = 56;
    b.y = 0;
    b.xv = 0;
    b.yv = 1;
    b.px = b.py = b.ppx = b.ppy = 0;
    b.pow = 0;
    b.land = 0;
    b.punch = 0;
    e.score[0] = e.score[1] = 0;
    e.round_ended = e.game_ended = e.p2serve = 0;
    e.has32 = 0;
    e.land_ok = 0;
    e.ep_frames = 0;
}

// ---- ball ------------------------------------------------------------------------------------
// process_collision_between_ball_and_world_and_set_ball_position, physics.py:359-436
__device__ __forceinline__ bool ball_world(Ball &b) {
    b.ppx = b.px;
    b.ppy = b.py;
    b.px = b.x;
    b.py = b.y;
    int fx = b.x + b.xv;
    if (fx < kBallRadius || fx > kGroundWidth) b.xv = -b.xv;  // asymmetric on purpose (:392-404)
    if (b.y + b.yv < 0) b.yv = 1;
    if (iabs(b.x - kGroundHalfWidth) < kNetHalfWidth && b.y > kNetTopTopY) {
        if (b.y <= kNetTopBottomY) {
            if (b.yv > 0) b.yv = -b.yv;
        } else {
            b.xv = (b.x < kGroundHalfWidth) ? -iabs(b.xv) : iabs(b.xv);
        }
    }
    int fy = b.y + b.yv;
    if (fy > kBallGroundY) {
        b.yv = -b.yv;
        b.punch = b.x;
        b.y = kBallGroundY;
        return true;
    }
    b.y = fy;
    b.x += b.xv;
    b.yv += 1;
    return false;
}

// Trajectory simulation shared by calculate_expected_landing_point_x_for (physics.py:643-686,
// POWER = false: net top test is the strict y < 192) and expected_landing_point_x_when_power_hit
// (physics.py:847-884, POWER = true: the whole net zone only bounces yv).
// Warp-collective over `mask`: every lane of `mask` must call it; lanes with active == false
// idle. Trip count 1..1000 per lane; the warp leaves when its slowest lane lands.
//
// Fast-forward: while the ball is in free flight the reference's loop body is the identity on
// everything but (x += xv, y += yv, yv += 1). kSkip iterations are applied in closed form when
// none of the loop's rules can fire in any of them:
//   wall    (x_j + xv outside [20,432], j < m)        <=> x_1 and x_m inside [20,432] (x is monotone)
//   ceiling (y_j + yv_j < 0, i.e. y_{j+1} < 0)        <=  yv >= 0 and y_1 >= 0 (y can be negative after a
//                                                         net bounce), or yv < 0 and y_0 - |yv|(|yv|+1)/2 >= 0
//   net     (|x_j - 216| < 25 and y_j > 176, j < m)   <=  y_0, y_{m-1} <= 176 (y is convex in j), or
//                                                         x_0 and x_{m-1} on one side outside (191,241)
//   ground  (y_{j+1} > 252)                           <=> y_1 and y_m <= 252 (convex)
//   limit   (loop_counter >= 1000)                    <=> it + m < 1000
// so the result (x at the break and whether the break was the ground) is unchanged.
constexpr int kSkip = 8;

template <bool POWER>
__device__ __forceinline__ int simulate_landing_x(unsigned mask, int x, int y, int xv, int yv, bool active,
                                                  bool &by_ground) {
    int it = 0;
    by_ground = false;
    while (__any_sync(mask, active)) {
        if (active) {
            constexpr int m = kSkip;
            const int x1 = x + xv, xm1 = x + (m - 1) * xv, xm = xm1 + xv;
            const int y1 = y + yv;
            const int ym1 = y + (m - 1) * yv + (m - 1) * (m - 2) / 2;
            const int ym = y + m * yv + m * (m - 1) / 2;
            const bool wall_ok = min(x1, xm) >= kBallRadius && max(x1, xm) <= kGroundWidth;
            const bool ceil_ok = yv >= 0 ? (y1 >= 0) : (2 * y >= yv * (yv - 1));
            const bool net_ok = max(y, ym1) <= kNetTopTopY ||
                                max(x, xm1) <= kGroundHalfWidth - kNetHalfWidth ||
                                min(x, xm1) >= kGroundHalfWidth + kNetHalfWidth;
            const bool ground_ok = max(y1, ym) <= kBallGroundY;
            if (wall_ok && ceil_ok && net_ok && ground_ok && it + m < kLoopLimit) {
                x = xm;
                y = ym;
                yv += m;
                it += m;
            } else {
                it += 1;
                if (x1 < kBallRadius || x1 > kGroundWidth) xv = -xv;
                if (y + yv < 0) yv = 1;
                if (iabs(x - kGroundHalfWidth) < kNetHalfWidth && y > kNetTopTopY) {
                    if (POWER) {
                        if (yv > 0) yv = -yv;
                    } else {
                        if (y < kNetTopBottomY) {
                            if (yv > 0) yv = -yv;
                        } else {
                            xv = (x < kGroundHalfWidth) ? -iabs(xv) : iabs(xv);
                        }
                    }
                }
                y += yv;
                if (y > kBallGroundY) {
                    active = false;
                    by_ground = true;
                } else if (it >= kLoopLimit) {
                    active = false;
                } else {
                    x += xv;
                    yv += 1;
                }
            }
        }
    }
    return x;
}

// expected landing x of the current ball for the lanes with `need`; memoised when possible.
template <bool TABLES = false>  // TABLES: the caller knows the memoised tables are there (one launch-uniform test less)
__device__ __forceinline__ void update_landing(unsigned mask, Env &e, const StepCfg &c, bool need) {
    const Ball &b = e.b;
    if (TABLES || c.tab_land != nullptr) {
        if (need && iabs(b.yv) <= kTabYv && iabs(b.xv) <= 20 && tab_pos_ok(b.x, b.y)) {
            const unsigned v = __ldg(c.tab_land + tab_land_index(b.x, b.y, b.xv, b.yv));
            e.b.land = (int)(v & 0x7FFFu);
            e.land_ok = (int)(v >> 15);
            need = false;
        }
    }
    if (__any_sync(mask, need)) {  // outside the memoised domain, or tables off
        bool g;
        const int lx = simulate_landing_x<false>(mask, b.x, b.y, b.xv, b.yv, need, g);
        if (need) {
            e.b.land = lx;
            e.land_ok = g;
        }
    }
}

// ---- computer player ---------------------------------------------------------------------------
// let_computer_decide_user_input (physics.py:689-771) + decide_whether_input_power_hit (:774-817).
// Warp-collective over `mask`.
template <int I, class Ctx, bool KFRAME = false, bool TABLES = false>
__device__ __forceinline__ void computer_decide(unsigned mask, Env &e, Ctx &d, const StepCfg &cfg, Input &in,
                                                int *scratch) {
    Player &p = e.p[I];
    const Player &o = e.p[1 - I];
    const Ball &b = e.b;
    constexpr int left_boundary = I * kGroundHalfWidth;
    constexpr int right_boundary = (I + 1) * kGroundHalfWidth;
    constexpr int far_boundary = I * kGroundWidth + kGroundHalfWidth;  // :718, :801

    in.xdir = 0;
    in.ydir = 0;
    in.power = 0;
    const int dx = iabs(b.x - p.x);

    int virt = b.land;
    if (dx > 100 && iabs(b.xv) < p.bold + 5) {
        if ((b.land <= left_boundary || b.land >= far_boundary) && p.standby == 0)
            virt = left_boundary + kGroundHalfWidth / 2;
    }
    const bool near = iabs(virt - p.x) <= p.bold + 8;
    if (!near) in.xdir = (p.x < virt) ? 1 : -1;

    bool search = false;
    if (p.state == 0) {
        if (iabs(b.xv) < p.bold + 3 && dx < kPlayerHalfLength && b.y > -36 && b.y < 10 * p.bold + 84 && b.yv > 0)
            in.ydir = -1;
        if (b.land > left_boundary && b.land < right_boundary && dx > p.bold * 5 + kPlayerLength &&
            b.x > left_boundary && b.x < right_boundary && b.y > 174) {
            in.power = 1;  // dive
            in.xdir = (p.x < b.x) ? 1 : -1;
        }
    } else if (p.state == 1 || p.state == 2) {
        if (dx > 8) in.xdir = (p.x < b.x) ? 1 : -1;
        search = dx < 48 && iabs(b.y - p.y) < 48;
    }

    // decide_whether_input_power_hit (:774-817): up to 6 candidate hits (x_direction 1,0 x three
    // y_directions), the first acceptable one in scan order wins. The candidate simulations have no
    // side effects, so all 6 are evaluated and then scanned in the reference's order: from the
    // memoised table when the ball is inside its domain, otherwise iteratively, spread over the
    // lanes of the warp (one (searcher, candidate) pair per lane and pass, through the warp's scratch).
    // Memoised power-hit candidates: six INDEPENDENT 2-byte loads (x_direction 1, 0 x half_yv0 = +H, 0, -H with
    // H = |ball yv|), issued by the searching lane itself BEFORE the draws: the draw (:795) only decides in which
    // order +H and -H are scanned, not which entries are read, so the ~80 instructions of the draws hide the loads'
    // latency. (Until round 2 the (searcher, candidate) pairs were spread over the lanes of the warp like the
    // iterative simulations below: five shuffles, a rank table in shared memory and three warp barriers per player
    // and frame cost more than the lanes that idle through these straight-line instructions.)
    const int ayv = iabs(b.yv);
    const bool tab_ok = (TABLES || cfg.tab_power != nullptr) && search && ayv <= kTabYv && tab_pos_ok(b.x, b.y);
    int lx_up[2] = {0, 0}, lx_mid[2] = {0, 0}, lx_dn[2] = {0, 0};  // [x_direction]
    int y_first = 0;
    auto load_candidates = [&]() {
        constexpr int kStrideH = kTabNy * kTabNx, kStrideXd = kTabNyv * kStrideH;
        const uint16_t *t0 = cfg.tab_power + tab_power_index(b.x, b.y, 0, 0);
        const uint16_t *t1 = t0 + kStrideXd;
        const int H = ayv * (KFRAME ? y_first : 1) * kStrideH;  // KFRAME: "up" is the side scanned first
        lx_up[1] = __ldg(t1 + H), lx_mid[1] = __ldg(t1), lx_dn[1] = __ldg(t1 - H);
        lx_up[0] = __ldg(t0 + H), lx_mid[0] = __ldg(t0), lx_dn[0] = __ldg(t0 - H);
    };
    // (the K-frame kernels hide latency with their six or seven CTAs per SM and are short of registers instead:
    // there the six values are not kept live across the draws, and the draw can then choose the addresses instead
    // of four selects — 1.82 against 1.84 ms per K = 64 launch)
    if (!KFRAME && tab_ok) load_candidates();

    // This frame's draws in the reference's order: :728 `integers(0, 20) == 0` and then :729 `integers(0, 2)`
    // if near, :795 `integers(0, 2)` (scan order of y_direction) if a power hit is searched for.
    int standby_draw, y_draw;
    d.computer_draws(e.has32, near, search, standby_draw, y_draw);
    if (standby_draw >= 0) p.standby = standby_draw;
    y_first = search ? (y_draw == 0 ? -1 : 1) : 0;
    bool found = false;
    if (tab_ok) {
        if (KFRAME) load_candidates();
        // acceptable (:806-811): lx <= left_boundary || lx >= far_boundary, and |lx - opponent x| > 64 — two unsigned
        // range tests (lx is never negative). Scan order: x_direction 1 then 0; y_direction y_first, 0, -y_first.
        constexpr unsigned side_lo = left_boundary + 1, side_span = far_boundary - left_boundary - 1;
        const bool up_first = KFRAME || y_first > 0;
        int first = -1;
#pragma unroll
        for (int c = 5; c >= 0; c--) {
            const int xd = c < 3 ? 1 : 0, r = c % 3;
            const int lx = r == 1 ? lx_mid[xd] : ((r == 0) == up_first ? lx_up[xd] : lx_dn[xd]);
            const bool ok = (unsigned)(lx - (int)side_lo) >= side_span &&
                            (unsigned)(lx - o.x + kPlayerLength) > 2u * kPlayerLength;
            first = ok ? c : first;  // the first acceptable candidate in the reference's scan order wins
        }
        if (first >= 0) {
            in.xdir = (first < 3) ? 1 : 0;
            in.ydir = y_first * (1 - (first < 3 ? first : first - 3));
            found = true;
        }
        search = false;
    }
    const unsigned sm = __ballot_sync(mask, search);
    if (sm) {  // warp-uniform: lanes outside the memoised domain, or tables off
        const int lane = threadIdx.x & 31;
        int4 *s_in = reinterpret_cast<int4 *>(scratch);  // [32] {ball x, ball y, |ball yv|, y_first}
        int *s_out = scratch + 128;                       // [32][6] landing x
        if (search) s_in[lane] = make_int4(b.x, b.y, ayv, y_first);
        __syncwarp(mask);
        const int total = 6 * __popc(sm), workers = __popc(mask);
        const int w = __popc(mask & ((1u << lane) - 1u));
#pragma unroll 1
        for (int base = 0; base < total; base += workers) {
            const int q = base + w;
            const bool act = q < total;
            const int r = act ? q / 6 : 0, c = act ? q - 6 * r : 0;
            const int leader = __fns(sm, 0, r + 1);  // lane of the r-th searcher
            const int4 in4 = s_in[leader];
            const int xd = (c < 3) ? 1 : 0;             // range(1, -1, -1)
            const int yd = in4.w * (1 - (c % 3));       // -1,0,1 or 1,0,-1
            const int xv0 = (in4.x < kGroundHalfWidth) ? (xd + 1) * 10 : -(xd + 1) * 10;  // :841-844
            const int yv0 = in4.z * yd * 2;                                                // :845
            bool g;
            const int lx = simulate_landing_x<true>(mask, in4.x, in4.y, xv0, yv0, act, g);
            if (act) s_out[leader * 6 + c] = lx;
        }
        __syncwarp(mask);
        if (search) {
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const int lx = s_out[lane * 6 + c];
                if (!found && (lx <= left_boundary || lx >= far_boundary) && iabs(lx - o.x) > kPlayerLength) {
                    in.xdir = (c < 3) ? 1 : 0;
                    in.ydir = y_first * (1 - (c % 3));
                    found = true;
                }
            }
        }
        __syncwarp(mask);  // scratch is reused by the other player's search
    }
    if (found) {  // :768-771
        in.power = 1;
        if (iabs(o.x - p.x) < 80 && in.ydir != -1) in.ydir = -1;
    }
}

// ---- player ------------------------------------------------------------------------------------
// The sprite animation at the end of process_player_movement_and_set_player_position (physics.py:524-552):
// (state, frame_number, delay_before_next_frame, normal_status_arm_swing_direction) -> the same. A pure function of
// 3 + 3 + 3 + 1 bits: the K-frame kernels read it from a 1,024-entry table in shared memory that they fill with this
// very function (anim_fill), most per-step kernels from the same table evaluated by the compiler (g_anim_table).
__host__ __device__ constexpr void player_animate(int &state, int &frame, int &delay, int &arm) {
    if (state == 1) {
        frame = frame >= 2 ? frame - 2 : frame + 1;  // (frame + 1) % 3 for frame in 0..4
    } else if (state == 2) {
        if (delay < 1) {
            frame += 1;
            if (frame > 4) {
                frame = 0;
                state = 1;
            }
        } else {
            delay -= 1;
        }
    } else if (state == 0) {
        delay += 1;
        if (delay > 3) {
            delay = 0;
            const int f = frame + arm;
            if (f < 0 || f > 4) arm = -arm;
            frame = frame + arm;
        }
    }
}
#ifndef PZ_STEP_ANIM_LUT
#define PZ_STEP_ANIM_LUT 1
#endif
constexpr int kAnimLutEntries = 1024;
__device__ __forceinline__ int anim_index(int state, int frame, int delay, int arm) {
    return state + frame * 8 + delay * 64 + (arm + 1) * 256;  // every field is at most 7 (pz_state.cuh), arm is +-1
}
// entry = state | frame << 8 | delay << 16 | (int8) arm << 24: a byte each, one PRMT each to take apart
__host__ __device__ constexpr uint32_t anim_entry(int index) {
    int state = index & 7, frame = (index >> 3) & 7, delay = (index >> 6) & 7, arm = (index & 512) ? 1 : -1;
    player_animate(state, frame, delay, arm);
    return (uint32_t)state | ((uint32_t)frame << 8) | ((uint32_t)delay << 16) | ((uint32_t)(arm & 0xFF) << 24);
}
__device__ __forceinline__ void anim_fill(uint32_t *lut, int tid, int nthreads) {  // caller synchronises
    for (int k = tid; k < kAnimLutEntries; k += nthreads) lut[k] = anim_entry(k);
}
// the same table evaluated by the compiler, in global memory, for the per-step kernels (read through L1)
struct AnimTable {
    uint32_t v[kAnimLutEntries];
};
__host__ __device__ constexpr AnimTable make_anim_table() {
    AnimTable t{};
    for (int k = 0; k < kAnimLutEntries; k++) t.v[k] = anim_entry(k);
    return t;
}
static __device__ const AnimTable g_anim_table = make_anim_table();

// process_player_movement_and_set_player_position, physics.py:439-564 (after the AI override). Written as selects:
// the lanes of a warp are in different phases of play, so every side of a branch runs anyway and the branches
// themselves (and the register moves at their joins) were a fifth of this function. ANIM_LUT: `anim` is the table
// above in shared memory.
template <int I, bool ANIM_LUT = false>
__device__ __forceinline__ void player_move(Player &p, const Input &in, const uint32_t *anim = nullptr) {
    if (p.state == 4) {  // lying down: don't move (:458-462)
        p.lying -= 1;
        if (p.lying < -1) p.state = 0;
        return;
    }
    constexpr int lo = I ? kGroundHalfWidth + kPlayerHalfLength : kPlayerHalfLength;
    constexpr int hi = I ? kGroundWidth - kPlayerHalfLength : kGroundHalfWidth - kPlayerHalfLength;
    const int vx = (p.state < 3) ? in.xdir * 6 : p.dive * 8;  // state < 5 always holds before termination
    p.x = min(max(p.x + vx, lo), hi);

    const bool jump = p.state < 3 && in.ydir == -1 && p.y == kPlayerGroundY;
    int state = jump ? 1 : p.state;
    int yv = jump ? -16 : p.yv;
    int frame = jump ? 0 : p.frame;
    const int fy = p.y + yv;  // gravity
    const bool air = fy < kPlayerGroundY, landing = fy > kPlayerGroundY;  // exactly on the ground: neither (:504-517)
    yv = air ? yv + 1 : (landing ? 0 : yv);
    p.y = landing ? kPlayerGroundY : fy;
    frame = landing ? 0 : frame;
    const bool dived = landing && state == 3;
    state = landing ? (dived ? 4 : 0) : state;
    p.lying = dived ? 3 : p.lying;
    const bool hit = in.power == 1 && state == 1;                     // power hit
    const bool dive = in.power == 1 && state == 0 && in.xdir != 0;    // dive
    int delay = hit ? 5 : p.delay;
    frame = (hit || dive) ? 0 : frame;
    state = hit ? 2 : (dive ? 3 : state);
    p.dive = dive ? in.xdir : p.dive;
    p.yv = dive ? -5 : yv;
    int arm = p.arm;
    if (ANIM_LUT) {
        const uint32_t v = anim[anim_index(state, frame, delay, arm)];
        state = (int)(v & 0xFFu);
        frame = (int)((v >> 8) & 0xFFu);
        delay = (int)((v >> 16) & 0xFFu);
        arm = (int)(int8_t)(v >> 24);
    } else {
        player_animate(state, frame, delay, arm);
    }
    p.state = state;
    p.frame = frame;
    p.delay = delay;
    p.arm = arm;
    // :554-564 (player.game_ended) cannot execute before the env terminates.
}

// is_collision_between_ball_and_player_happened (physics.py:340-356) and
// process_collision_between_ball_and_player (physics.py:580-640). Returns true on a NEW collision.
template <int I, class Ctx>
__device__ __forceinline__ bool ball_player(Env &e, Ctx &d, const Input &in) {
    Player &p = e.p[I];
    Ball &b = e.b;
    // |dx| <= 32 and |dy| <= 32 (:340-356) as two unsigned range tests
    const bool hit = (unsigned)(b.x - p.x + kPlayerHalfLength) <= 2u * kPlayerHalfLength &&
                     (unsigned)(b.y - p.y + kPlayerHalfLength) <= 2u * kPlayerHalfLength;
    if (!hit) {
        p.coll = 0;
        return false;
    }
    if (p.coll) return false;
    p.coll = 1;
    if (b.x < p.x)
        b.xv = -(iabs(b.x - p.x) / 3);
    else if (b.x > p.x)
        b.xv = iabs(b.x - p.x) / 3;
    if (b.xv == 0) b.xv = d.template integers<3>(e.has32) - 1;  // :613
    const int ayv = iabs(b.yv);
    b.yv = (ayv < 15) ? -15 : -ayv;
    if (p.state == 2) {  // jumping and power hitting
        b.xv = (b.x < kGroundHalfWidth) ? (iabs(in.xdir) + 1) * 10 : -(iabs(in.xdir) + 1) * 10;
        b.punch = b.x;
        b.yv = iabs(b.yv) * in.ydir * 2;
        b.pow = 1;
    } else {
        b.pow = 0;
    }
    return true;
}

// ---- one frame -----------------------------------------------------------------------------------
// raw_env.step (pikazoo_env.py:175-240) for an env that has not terminated. keys1/keys2 are the
// decoded key bits. AI_MASK bit I = player I+1 is a computer. Warp-collective over `mask` when
// AI_MASK != 0. Returns player_1's base reward (-1, 0, +1).
// The same with the two players' inputs already decoded (get_input has run: it only touches
// power_hit_key_is_down_previous, which the new-round block below does not, so the order is immaterial).
// KFRAME (the K-frame kernels): `anim` is the sprite-animation table in shared memory (anim_fill), and the computer
// players' table loads are not hoisted above the draws.
template <int AI_MASK, class Ctx, bool KFRAME = false, bool ANIM_LUT = KFRAME, bool TABLES = false>
__device__ __forceinline__ int step_frame_inputs(unsigned mask, Env &e, Ctx &d, const StepCfg &c, Input in1, Input in2,
                                                 int *scratch, const uint32_t *anim = nullptr) {
    if (e.round_ended) {  // :176-180 (game_ended is false here)
        new_round(e, d, c);
        e.round_ended = 0;
    }

    // physics_engine, physics.py:280-337
    // The real ball update differs from one iteration of the landing simulation only on the ground
    // (touching) and for y == 192 inside the net zone (<= vs <, :412 vs :670).
    const bool net192 = e.b.y == kNetTopBottomY && iabs(e.b.x - kGroundHalfWidth) < kNetHalfWidth;
    const bool touching = ball_world(e.b);
    if (AI_MASK != 0) {
        // :314-315 is evaluated twice per frame on an unchanged ball; once is enough. And while the
        // ball free-flies along the trajectory that was simulated last frame (land_ok: that
        // simulation ended on the ground), the landing point of the advanced ball is the same.
        update_landing<TABLES>(mask, e, c, !e.land_ok || touching || net192);
    } else {
        e.land_ok = 0;  // expected_landing_point_x is not maintained without computer players
    }
    if (AI_MASK & 1) computer_decide<0, Ctx, KFRAME, TABLES>(mask, e, d, c, in1, scratch);
    player_move<0, ANIM_LUT>(e.p[0], in1, anim);
    if (AI_MASK & 2) computer_decide<1, Ctx, KFRAME, TABLES>(mask, e, d, c, in2, scratch);
    player_move<1, ANIM_LUT>(e.p[1], in2, anim);

    bool recalc = ball_player<0>(e, d, in1);
    recalc |= ball_player<1>(e, d, in2);
    if (AI_MASK != 0) {
        // :331-332 after each new collision; only the value for the final ball state survives.
        if (__any_sync(mask, recalc)) update_landing<TABLES>(mask, e, c, recalc);
    }

    // scoring, :190-210
    if (touching) {
        if (e.b.punch < kGroundHalfWidth) {
            e.p2serve = 1;
            e.score[1] += 1;
            if (e.score[1] >= c.winning_score) e.game_ended = 1;
        } else {
            e.p2serve = 0;
            e.score[0] += 1;
            if (e.score[0] >= c.winning_score) e.game_ended = 1;
        }
        e.round_ended = 1;
    }
    e.ep_frames += 1;
    return e.round_ended ? (e.p2serve ? -1 : 1) : 0;  // :217-223
}

template <int AI_MASK, class Ctx>
__device__ __forceinline__ int step_frame(unsigned mask, Env &e, Ctx &d, const StepCfg &c, uint32_t keys1,
                                          uint32_t keys2, int *scratch) {
    const Input in1 = get_input(e.p[0], keys1);  // also runs for computer players (:183-184)
    const Input in2 = get_input(e.p[1], keys2);
    return step_frame_inputs<AI_MASK>(mask, e, d, c, in1, in2, scratch);
}

// get_input from a pre-decoded action: (x_direction + 1) | (y_direction + 1) << 2 | power key << 4
__host__ __device__ constexpr uint32_t pack_input(uint32_t keys) {
    return (uint32_t)(((keys & kL) ? -1 : ((keys & kR) ? 1 : 0)) + 1) |
           ((uint32_t)(((keys & kU) ? -1 : ((keys & kD) ? 1 : 0)) + 1) << 2) | ((keys & kP) ? 16u : 0u);
}
__device__ __forceinline__ Input input_from_packed(Player &p, uint32_t q) {
    Input in;
    in.xdir = (int)(q & 3u) - 1;
    in.ydir = (int)((q >> 2) & 3u) - 1;
    const int down = (int)(q >> 4);
    in.power = down & (p.keyprev ^ 1);
    p.keyprev = down;
    return in;
}

// raw_env._get_obs, pikazoo_env.py:576-624: the 35 distinct values (p1 block 13, p2 block 13,
// ball block 9); obs_p1 = u[0..34], obs_p2 = u[13..25] u[0..12] u[26..34].
__device__ __forceinline__ void obs_values(const Env &e, int (&u)[35]) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const Player &p = e.p[i];
        int *o = u + 13 * i;
        o[0] = p.x;
        o[1] = p.y;
        o[2] = p.yv;
        o[3] = p.dive;
        o[4] = p.lying;
        o[5] = p.frame;
        o[6] = p.delay;
#pragma unroll
        for (int s = 0; s < 5; s++) o[7 + s] = (p.state == s) ? 1 : 0;
        o[12] = p.keyprev;
    }
    const Ball &b = e.b;
    u[26] = b.x;
    u[27] = b.y;
    u[28] = b.px;
    u[29] = b.py;
    u[30] = b.ppx;
    u[31] = b.ppy;
    u[32] = b.xv;
    u[33] = b.yv;
    u[34] = b.pow;
}

// NormalizeObservation (pikazoo/wrappers/normalize_observation.py:18-32): (obs - low) / (high - low)
// with the bounds of raw_env.observation_space (pikazoo_env.py:485-562), per distinct value u[k].
// The reference divides int64 arrays with numpy true division, i.e. one correctly rounded double
// division per element; the float32 form is float32(that double), which equals the correctly rounded
// float32 division of the same two integers (double rounding is innocuous for p = 24 into P = 53 >= 2p + 2).
__host__ __device__ constexpr int obs_low(int k) {
    return k < 26 ? ((k % 13) == 0 ? kPlayerHalfLength : (k % 13) == 1 ? 108 : (k % 13) == 2 ? -15
                     : (k % 13) == 3 ? -1 : (k % 13) == 4 ? -2 : 0)
                  : (k == 26 ? kBallRadius : k == 32 ? -20 : k == 33 ? -124 : 0);
}
__host__ __device__ constexpr int obs_high(int k) {
    return k < 26 ? ((k % 13) == 0 ? kGroundWidth - kPlayerHalfLength : (k % 13) == 1 ? kPlayerGroundY
                     : (k % 13) == 2 ? 16 : (k % 13) == 3 ? 1 : (k % 13) == 4 ? 3 : ((k % 13) < 7 ? 4 : 1))
                  : ((k == 26 || k == 28 || k == 30) ? kGroundWidth
                     : (k == 27 || k == 29 || k == 31) ? kBallGroundY : k == 32 ? 20 : k == 33 ? 124 : 1);
}

// float32(n / D) for a compile-time divisor without a division: q = n * RN(1/D), one FMA for the exact
// residual n - q*D, one FMA to apply it (Markstein's correction; correctly rounded for every numerator
// the packed state can hold — checked exhaustively on the host by tests/test_device_code_on_host.py).
// Powers of two multiply exactly.
template <int D>
__device__ __forceinline__ float div_const_f32(int n) {
    const float a = (float)n;  // exact: |n| < 2^24
    if ((D & (D - 1)) == 0) return a * (1.0f / (float)D);
    constexpr float r = 1.0f / (float)D;
    const float q = a * r;
    const float rem = fmaf(-q, (float)D, a);
    return fmaf(rem, r, q);
}

// element K of the distinct values, as floating point
template <typename F, int K>
__device__ __forceinline__ F obs_float(const int (&u)[35], bool normalize) {
    if (!normalize) return (F)u[K];
    if (sizeof(F) == 4) return (F)div_const_f32<obs_high(K) - obs_low(K)>(u[K] - obs_low(K));
    return (F)(u[K] - obs_low(K)) / (F)(obs_high(K) - obs_low(K));
}

template <typename F, int... K>
__device__ __forceinline__ void obs_floats_impl(const int (&u)[35], F (&f)[35], bool normalize,
                                                std::integer_sequence<int, K...>) {
    ((f[K] = obs_float<F, K>(u, normalize)), ...);
}
// all 35 distinct values (the two agents' rows are permutations of them)
template <typename F>
__device__ __forceinline__ void obs_floats(const int (&u)[35], F (&f)[35], bool normalize) {
    obs_floats_impl(u, f, normalize, std::make_integer_sequence<int, 35>{});
}

// index into u[] of element k (0..69) of the [obs_p1 | obs_p2] row
__host__ __device__ constexpr int obs_src(int k) {
    return k < 35 ? k : (k < 48 ? k - 35 + 13 : (k < 61 ? k - 48 : k - 35));
}

}  // namespace pz
