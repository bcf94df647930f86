// Packed per-env device state (17 int32 words, structure-of-arrays in 16-byte groups) and
// its register-resident form. See DESIGN.md §3 for the layout rationale.
//
// Field inventory follows the reference objects:
//   Player        pikazoo/env/physics.py:140-218   (+ PikaUserInput.power_hit_key_is_down_previous :51)
//   Ball          pikazoo/env/physics.py:221-277   (render-only fields dropped)
//   raw_env       pikazoo/env/pikazoo_env.py:100-111 (scores, round_ended, game_ended, is_player2_serve)
//   PCG64         numpy bit generator: 128-bit state, 128-bit inc, has_uint32, uinteger
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pz {

// ---- physics constants (pikazoo/env/physics.py:9-33) ------------------------------------
constexpr int kGroundWidth = 432;
constexpr int kGroundHalfWidth = 216;
constexpr int kPlayerLength = 64;
constexpr int kPlayerHalfLength = 32;
constexpr int kPlayerGroundY = 244;
constexpr int kBallRadius = 20;
constexpr int kBallGroundY = 252;
constexpr int kNetHalfWidth = 25;
constexpr int kNetTopTopY = 176;
constexpr int kNetTopBottomY = 192;
constexpr int kLoopLimit = 1000;

// ---- register-resident env ---------------------------------------------------------------
struct Player {
    int x, y, yv, state, frame, delay, arm, dive, lying, coll, bold, standby, keyprev;
};
struct Ball {
    int x, y, xv, yv, px, py, ppx, ppy, pow, land, punch;
};
struct Env {
    Player p[2];
    Ball b;
    int score[2];
    int round_ended, game_ended, p2serve;
    int has32;       // PCG64 buffered-half flag (lives with the env word so it is always resident)
    int land_ok;     // derived: b.land == landing x of the CURRENT ball and that simulation ended on the
                     // ground (not at the loop limit), so it stays valid while the ball free-flies
    int ep_frames;   // step() calls since reset (RecordEpisodeStatistics length)
};

// PCG64 stream of one env; loaded lazily by the per-step kernels (most frames draw nothing).
struct Rng {
    uint64_t s_lo, s_hi;     // 128-bit LCG state
    uint64_t inc_lo, inc_hi; // 128-bit increment (read-only after seeding)
    uint32_t uinteger;       // buffered high half of the last next64
    bool loaded, dirty;
};

// ---- global-memory layout ------------------------------------------------------------------
// state_dev is int32[17*N]:  G0 int4[N] | G1 int4[N] | G2 int4[N] | G3 int4[N] | U uint32[N]
//   G0 = {P1 bits 0..31, P2 bits 0..31, P1 bits 32..43 | P2 bits 32..43 << 12, ep_frames}
//   G1 = {B0, B1, B2, ENV}
//   G2 = PCG64 state (little-endian 32-bit words), G3 = PCG64 inc, U = uinteger
// L2 eviction policies (the 64-bit operand of `.L2::cache_hint`; values of createpolicy.fractional
// with fraction 1.0). Observations / rewards / dones are written once and never read back by the
// simulator, so their stores ask to be evicted first; state accesses use the normal policy (see
// fill_params in pz_kernels.cu for the measurements behind both choices).
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ULL;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ULL;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ULL;

struct StatePtrs {
    int4 *g0, *g1, *g2, *g3;
    uint32_t *u;
    uint64_t policy;  // L2 cache policy of every state access
};

__host__ __device__ inline StatePtrs state_ptrs(int32_t *base, int64_t n, uint64_t policy = kL2EvictNormal) {
    StatePtrs s;
    s.policy = policy;
    s.g0 = reinterpret_cast<int4 *>(base);
    s.g1 = s.g0 + n;
    s.g2 = s.g1 + n;
    s.g3 = s.g2 + n;
    s.u = reinterpret_cast<uint32_t *>(s.g3 + n);
    return s;
}

// Player, 44 bits:
//   x:9 | y:8 | (yv+32):6 | state:3 | frame:3 | delay:3 || arm(+1 -> 1):1 | (dive+1):2 |
//   (lying+4):3 | coll:1 | bold:3 | standby:1 | keyprev:1
__device__ __forceinline__ uint32_t pack_player_lo(const Player &p) {
    return (uint32_t)p.x | ((uint32_t)p.y << 9) | ((uint32_t)(p.yv + 32) << 17) | ((uint32_t)p.state << 23) |
           ((uint32_t)p.frame << 26) | ((uint32_t)p.delay << 29);
}
__device__ __forceinline__ uint32_t pack_player_hi(const Player &p) {
    return (uint32_t)(p.arm > 0) | ((uint32_t)(p.dive + 1) << 1) | ((uint32_t)(p.lying + 4) << 3) |
           ((uint32_t)p.coll << 6) | ((uint32_t)p.bold << 7) | ((uint32_t)p.standby << 10) |
           ((uint32_t)p.keyprev << 11);
}
__device__ __forceinline__ void unpack_player(Player &p, uint32_t lo, uint32_t hi) {
    p.x = lo & 511;
    p.y = (lo >> 9) & 255;
    p.yv = (int)((lo >> 17) & 63) - 32;
    p.state = (lo >> 23) & 7;
    p.frame = (lo >> 26) & 7;
    p.delay = (lo >> 29) & 7;
    p.arm = (hi & 1) ? 1 : -1;
    p.dive = (int)((hi >> 1) & 3) - 1;
    p.lying = (int)((hi >> 3) & 7) - 4;
    p.coll = (hi >> 6) & 1;
    p.bold = (hi >> 7) & 7;
    p.standby = (hi >> 10) & 1;
    p.keyprev = (hi >> 11) & 1;
}

// Ball + env, 4 words:
//   B0  = yv:16 (two's complement) | (xv+32):6 << 16 | pow << 22 | land:9 << 23
//   B1  = x:9 | px:9 << 9 | ppx:9 << 18 | has_uint32 << 27 | land_ok << 28
//   B2  = (y+512):10 | (py+512):10 << 10 | (ppy+512):10 << 20
//   ENV = score1:10 | score2:10 << 10 | round_ended << 20 | game_ended << 21 | p2serve << 22 | punch:9 << 23
// Ball x is unsigned: the wall rule (physics.py:392-404) keeps it in [20, 432] from either serve position
// (a flip turns x + xv < 20 into x + |xv| > x). Ball y is signed: with |yv| > 176 the net-top bounce (physics.py:412-414, applied after the
// ceiling test :408-409) throws the ball above the ceiling for one frame, y = 177 - |yv| at the
// lowest; 10 bits cover |yv| <= 689, 16 bits of yv cover every velocity the doubling power hits
// can build before the ball leaves the court.
__device__ __forceinline__ int4 pack_g1(const Env &e) {
    const Ball &b = e.b;
    int4 w;
    w.x = (int)(((uint32_t)b.yv & 0xFFFFu) | ((uint32_t)(b.xv + 32) << 16) | ((uint32_t)b.pow << 22) |
                ((uint32_t)b.land << 23));
    w.y = (int)((uint32_t)b.x | ((uint32_t)b.px << 9) | ((uint32_t)b.ppx << 18) | ((uint32_t)e.has32 << 27) |
                ((uint32_t)e.land_ok << 28));
    w.z = (int)((uint32_t)(b.y + 512) | ((uint32_t)(b.py + 512) << 10) | ((uint32_t)(b.ppy + 512) << 20));
    w.w = (int)((uint32_t)e.score[0] | ((uint32_t)e.score[1] << 10) | ((uint32_t)e.round_ended << 20) |
                ((uint32_t)e.game_ended << 21) | ((uint32_t)e.p2serve << 22) | ((uint32_t)b.punch << 23));
    return w;
}
__device__ __forceinline__ void unpack_g1(Env &e, int4 w) {
    Ball &b = e.b;
    uint32_t b0 = (uint32_t)w.x, b1 = (uint32_t)w.y, b2 = (uint32_t)w.z, ev = (uint32_t)w.w;
    b.yv = (int)(int16_t)(b0 & 0xFFFFu);
    b.xv = (int)((b0 >> 16) & 63) - 32;
    b.pow = (b0 >> 22) & 1;
    b.land = (b0 >> 23) & 511;
    b.x = b1 & 511;
    b.px = (b1 >> 9) & 511;
    b.ppx = (b1 >> 18) & 511;
    e.has32 = (b1 >> 27) & 1;
    e.land_ok = (b1 >> 28) & 1;
    b.y = (int)(b2 & 1023) - 512;
    b.py = (int)((b2 >> 10) & 1023) - 512;
    b.ppy = (int)((b2 >> 20) & 1023) - 512;
    e.score[0] = ev & 1023;
    e.score[1] = (ev >> 10) & 1023;
    e.round_ended = (ev >> 20) & 1;
    e.game_ended = (ev >> 21) & 1;
    e.p2serve = (ev >> 22) & 1;
    b.punch = (ev >> 23) & 511;
}

__device__ __forceinline__ int4 pack_g0(const Env &e) {
    int4 w;
    w.x = (int)pack_player_lo(e.p[0]);
    w.y = (int)pack_player_lo(e.p[1]);
    w.z = (int)(pack_player_hi(e.p[0]) | (pack_player_hi(e.p[1]) << 12));
    w.w = e.ep_frames;
    return w;
}
__device__ __forceinline__ void unpack_g0(Env &e, int4 w) {
    uint32_t hi = (uint32_t)w.z;
    unpack_player(e.p[0], (uint32_t)w.x, hi & 0xFFFu);
    unpack_player(e.p[1], (uint32_t)w.y, (hi >> 12) & 0xFFFu);
    e.ep_frames = w.w;
}

// 128-bit state accesses: touched once per launch (keep them out of L1), L2 policy from the caller.
__device__ __forceinline__ int4 ld_stream(const int4 *p, uint64_t policy) {
#ifdef PZ_HOST_EMULATION  // tests/emul: the device code compiled for the host, one lane per warp
    (void)policy;
    return *p;
#else
    int4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
#endif
}
__device__ __forceinline__ void st_stream(int4 *p, int4 v, uint64_t policy) {
#ifdef PZ_HOST_EMULATION
    (void)policy;
    *p = v;
#else
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
                 : "memory");
#endif
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p, uint64_t policy) {
#ifdef PZ_HOST_EMULATION
    (void)policy;
    return *p;
#else
    uint32_t r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
    return r;
#endif
}
__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v, uint64_t policy) {
#ifdef PZ_HOST_EMULATION
    (void)policy;
    *p = v;
#else
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
#endif
}

__device__ __forceinline__ void load_env(Env &e, const StatePtrs &s, int64_t i) {
    int4 a = ld_stream(s.g0 + i, s.policy);
    int4 b = ld_stream(s.g1 + i, s.policy);
    unpack_g0(e, a);
    unpack_g1(e, b);
}
__device__ __forceinline__ void store_env(const Env &e, const StatePtrs &s, int64_t i) {
    st_stream(s.g0 + i, pack_g0(e), s.policy);
    st_stream(s.g1 + i, pack_g1(e), s.policy);
}

__device__ __forceinline__ void rng_load(Rng &r, const StatePtrs &s, int64_t i) {
    int4 st = ld_stream(s.g2 + i, s.policy);
    int4 ic = ld_stream(s.g3 + i, s.policy);
    r.s_lo = (uint64_t)(uint32_t)st.x | ((uint64_t)(uint32_t)st.y << 32);
    r.s_hi = (uint64_t)(uint32_t)st.z | ((uint64_t)(uint32_t)st.w << 32);
    r.inc_lo = (uint64_t)(uint32_t)ic.x | ((uint64_t)(uint32_t)ic.y << 32);
    r.inc_hi = (uint64_t)(uint32_t)ic.z | ((uint64_t)(uint32_t)ic.w << 32);
    r.uinteger = ld_stream_u32(s.u + i, s.policy);
    r.loaded = true;
}
__device__ __forceinline__ void rng_store(const Rng &r, const StatePtrs &s, int64_t i) {
    int4 st;
    st.x = (int)(uint32_t)r.s_lo;
    st.y = (int)(uint32_t)(r.s_lo >> 32);
    st.z = (int)(uint32_t)r.s_hi;
    st.w = (int)(uint32_t)(r.s_hi >> 32);
    st_stream(s.g2 + i, st, s.policy);
    st_stream_u32(s.u + i, r.uinteger, s.policy);
}

// ---- unpacked parity form (int32[53], oracle/pika_oracle.h pk_env; words 42..51 are the PCG64
// stream and are handled by the callers) -------------------------------------------------------
__device__ __forceinline__ void env_to_unpacked(const Env &e, int32_t *o) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const Player &p = e.p[k];
        int32_t *q = o + 13 * k;
        q[0] = p.x, q[1] = p.y, q[2] = p.yv, q[3] = p.state, q[4] = p.frame, q[5] = p.delay, q[6] = p.arm;
        q[7] = p.dive, q[8] = p.lying, q[9] = p.coll, q[10] = p.bold, q[11] = p.standby, q[12] = p.keyprev;
    }
    const Ball &b = e.b;
    int32_t *q = o + 26;
    q[0] = b.x, q[1] = b.y, q[2] = b.xv, q[3] = b.yv, q[4] = b.px, q[5] = b.py, q[6] = b.ppx, q[7] = b.ppy;
    q[8] = b.pow, q[9] = b.land, q[10] = b.punch;
    o[37] = e.score[0], o[38] = e.score[1], o[39] = e.round_ended, o[40] = e.game_ended, o[41] = e.p2serve;
    o[50] = e.has32;
    o[52] = e.ep_frames;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Values outside the packed field ranges are clamped (they cannot occur in a state produced by
// the simulator itself). The derived landing-cache bit is cleared.
__device__ __forceinline__ void env_from_unpacked(Env &e, const int32_t *o) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        Player &p = e.p[k];
        const int32_t *q = o + 13 * k;
        p.x = clampi(q[0], 0, 511), p.y = clampi(q[1], 0, 255), p.yv = clampi(q[2], -32, 31);
        p.state = clampi(q[3], 0, 7), p.frame = clampi(q[4], 0, 7), p.delay = clampi(q[5], 0, 7);
        p.arm = q[6] > 0 ? 1 : -1, p.dive = clampi(q[7], -1, 1), p.lying = clampi(q[8], -4, 3);
        p.coll = q[9] != 0, p.bold = clampi(q[10], 0, 7), p.standby = q[11] != 0, p.keyprev = q[12] != 0;
    }
    Ball &b = e.b;
    const int32_t *q = o + 26;
    b.x = clampi(q[0], 0, 511), b.y = clampi(q[1], -512, 511), b.xv = clampi(q[2], -32, 31);
    b.yv = clampi(q[3], -32768, 32767), b.px = clampi(q[4], 0, 511), b.py = clampi(q[5], -512, 511);
    b.ppx = clampi(q[6], 0, 511), b.ppy = clampi(q[7], -512, 511), b.pow = q[8] != 0;
    b.land = clampi(q[9], 0, 511), b.punch = clampi(q[10], 0, 511);
    e.score[0] = clampi(o[37], 0, 1023), e.score[1] = clampi(o[38], 0, 1023);
    e.round_ended = o[39] != 0, e.game_ended = o[40] != 0, e.p2serve = o[41] != 0;
    e.has32 = o[50] != 0;
    e.land_ok = 0;
    e.ep_frames = o[52];
}

}  // namespace pz
