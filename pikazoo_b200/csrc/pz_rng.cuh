// Device restatement of the numpy random stream the reference draws from
// (gymnasium.utils.seeding.np_random -> numpy Generator(PCG64), pikazoo_env.py:570-571).
// Third-party arithmetic, absent from /root/reference; algorithm per numpy 2.3.5
// (bit_generator.pyx SeedSequence, src/pcg64/pcg64.h, src/distributions/distributions.c)
// as restated and verified in SURVEY.md §8(c). 128-bit arithmetic is decomposed into
// 64-bit mul / mul.hi, which is what sm_100a executes natively.
#pragma once
#include "pz_state.cuh"

namespace pz {

constexpr uint64_t kPcgMultHi = 0x2360ED051FC65DA4ULL;
constexpr uint64_t kPcgMultLo = 0x4385DF649FCCF645ULL;

// state = state * MULT + inc (mod 2^128)
__device__ __forceinline__ void pcg_step(uint64_t &lo, uint64_t &hi, uint64_t inc_lo, uint64_t inc_hi) {
    uint64_t m_lo = lo * kPcgMultLo;
    uint64_t m_hi = __umul64hi(lo, kPcgMultLo) + hi * kPcgMultLo + lo * kPcgMultHi;
    uint64_t n_lo = m_lo + inc_lo;
    uint64_t carry = (n_lo < m_lo) ? 1ULL : 0ULL;
    lo = n_lo;
    hi = m_hi + inc_hi + carry;
}

// rotr64(x, rot), rot in [0, 63]: a rotation by 32 is a swap of the halves, the rest two funnel shifts (five
// instructions against the nine of the shift-and-or form)
__device__ __forceinline__ uint64_t rotr64(uint64_t x, unsigned rot) {
#ifdef PZ_HOST_EMULATION
    return (x >> rot) | (x << ((64u - rot) & 63u));
#else
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    const bool swap = (rot & 32u) != 0u;
    const uint32_t a = swap ? xh : xl, b = swap ? xl : xh;
    return (uint64_t)__funnelshift_r(a, b, rot) | ((uint64_t)__funnelshift_r(b, a, rot) << 32);
#endif
}

// pcg64 "setseq XSL-RR 128/64": advance, then output rotr64(hi ^ lo, hi >> 58)
__device__ __forceinline__ uint64_t pcg_next64(Rng &r) {
    pcg_step(r.s_lo, r.s_hi, r.inc_lo, r.inc_hi);
    uint64_t x = r.s_hi ^ r.s_lo;
    return rotr64(x, (unsigned)(r.s_hi >> 58));
}

// numpy pcg64_next32: low half first, high half buffered across calls (part of env state)
__device__ __forceinline__ uint32_t pcg_next32(Rng &r, int &has32) {
    r.dirty = true;
    if (has32) {
        has32 = 0;
        return r.uinteger;
    }
    uint64_t n = pcg_next64(r);
    has32 = 1;
    r.uinteger = (uint32_t)(n >> 32);
    return (uint32_t)n;
}

// Generator.integers(0, HIGH) scalar path: buffered_bounded_lemire_uint32 without a local
// buffer. HIGH is a compile-time constant at every reference call site
// (physics.py:218 -> 5, :613 -> 3, :728 -> 20, :729/:795 and pikazoo_env.py:246 -> 2).
template <uint32_t HIGH>
__device__ __forceinline__ int rng_integers(Rng &r, int &has32) {
    constexpr uint32_t rng_excl = HIGH;
    constexpr uint32_t threshold = (0xFFFFFFFFu - (HIGH - 1u)) % HIGH;
    uint64_t m = (uint64_t)pcg_next32(r, has32) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        while (threshold != 0u && leftover < threshold) {  // threshold == 0 for HIGH = 2
            m = (uint64_t)pcg_next32(r, has32) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (int)(m >> 32);
}

// Fused call sites. The lanes of a warp are uncorrelated in has32, so the branch inside pcg_next32 makes a
// warp run the generator step AND the buffered path at every draw, once per draw of a chain. Two draws that
// follow each other consume exactly two 32-bit halves, i.e. at most ONE new 64-bit output whatever has32 is:
// the step is taken speculatively once, the halves are selected, and the stream state is committed by
// selects. Lemire's rejection (leftover < HIGH, p < 5e-9 per draw) leaves through the general path from
// the untouched stream, so the values and the stream are those of the reference's calls, always.

// All draws of one computer player in one frame (let_computer_decide_user_input, physics.py:689-771); which of
// them happen is known before the first one:
//   near   (within reach of the target, :727):  a = integers(0, 20) (:728); if a == 0: standby = integers(0, 2) (:729)
//   search (a power hit is considered, :766):   y_first = integers(0, 2) (decide_whether_input_power_hit, :795)
// standby = -1 when :729 did not run. Two halves at most are served from the one speculative step; the
// 1-in-20 frame that needs three takes the general path, like the rejection corner.
__device__ __forceinline__ void rng_computer_draws(Rng &r, int &has32, bool near, bool search, int &standby,
                                                   int &y_first) {
    uint64_t lo = r.s_lo, hi = r.s_hi;
    pcg_step(lo, hi, r.inc_lo, r.inc_hi);
    const uint64_t x = hi ^ lo;
    const uint64_t n = rotr64(x, (unsigned)(hi >> 58));
    const bool buffered = has32 != 0;
    const uint32_t h0 = buffered ? r.uinteger : (uint32_t)n;                // the next two 32-bit halves
    const uint32_t h1 = buffered ? (uint32_t)n : (uint32_t)(n >> 32);
    const uint64_t m1 = (uint64_t)h0 * 20u;
    const bool second = near && (uint32_t)(m1 >> 32) == 0u;
    // halves consumed: k = near + second + search, kept as predicates (second implies near)
    if ((near && (uint32_t)m1 < 20u) || (second && search)) {  // rejection test applies, or k == 3: general path, from the untouched stream
        standby = -1;
        if (near && rng_integers<20>(r, has32) == 0) standby = rng_integers<2>(r, has32);
        if (search) y_first = rng_integers<2>(r, has32);
        return;
    }
    const bool any = near || search;                // k != 0
    const bool two = near && (second || search);    // k == 2 (k == 3 has left above)
    standby = second ? (int)(h1 >> 31) : -1;       // integers(0, 2) = (v * 2) >> 32, never rejects (threshold 0)
    y_first = (int)((near ? h1 : h0) >> 31);       // meaningful only if search (then `second` is false)
    if (buffered ? two : any) {                     // the new output was consumed (at least its low half)
        r.s_lo = lo;
        r.s_hi = hi;
        r.uinteger = (uint32_t)(n >> 32);
    }
    has32 = (buffered ? 1 : 0) ^ ((any && !two) ? 1 : 0);  // k & 1
    if (any) r.dirty = true;
}

// `a = integers(0, 5); b = integers(0, 5)`  (physics.py:218 for player 1 then player 2)
__device__ __forceinline__ void rng_integers5_twice(Rng &r, int &has32, int &a, int &b) {
    uint64_t lo = r.s_lo, hi = r.s_hi;
    pcg_step(lo, hi, r.inc_lo, r.inc_hi);
    const uint64_t x = hi ^ lo;
    const uint64_t n = rotr64(x, (unsigned)(hi >> 58));
    const bool buffered = has32 != 0;
    const uint64_t m1 = (uint64_t)(buffered ? r.uinteger : (uint32_t)n) * 5u;
    const uint64_t m2 = (uint64_t)(buffered ? (uint32_t)n : (uint32_t)(n >> 32)) * 5u;
    if ((uint32_t)m1 < 5u || (uint32_t)m2 < 5u) {  // the rejection test applies: general path
        a = rng_integers<5>(r, has32);
        b = rng_integers<5>(r, has32);
        return;
    }
    // buffered: the second draw stepped and buffered the new high half; otherwise the first did and the
    // second consumed it — has32 ends where it started, uinteger holds the new high half either way
    r.s_lo = lo;
    r.s_hi = hi;
    r.uinteger = (uint32_t)(n >> 32);
    r.dirty = true;
    a = (int)(m1 >> 32);
    b = (int)(m2 >> 32);
}

// ---- SeedSequence(seed) -> PCG64 state/inc (numpy bit_generator.pyx) ----------------------
__device__ __forceinline__ uint32_t ss_hashmix(uint32_t value, uint32_t &hash_const) {
    value ^= hash_const;
    hash_const *= 0x931e8875u;
    value *= hash_const;
    value ^= value >> 16;
    return value;
}
__device__ __forceinline__ uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    r ^= r >> 16;
    return r;
}

// 128-bit helpers on (lo, hi) pairs
__device__ __forceinline__ void u128_add(uint64_t &lo, uint64_t &hi, uint64_t a_lo, uint64_t a_hi) {
    uint64_t n = lo + a_lo;
    hi = hi + a_hi + ((n < lo) ? 1ULL : 0ULL);
    lo = n;
}

__device__ inline void pcg64_seed(uint64_t seed, Rng &r) {
    uint32_t entropy[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_entropy = entropy[1] ? 2 : 1;  // minimal little-endian uint32 split (0 -> [0])
    uint32_t pool[4];
    uint32_t hc = 0x43b0d7e5u;
#pragma unroll
    for (int i = 0; i < 4; i++) pool[i] = ss_hashmix(i < n_entropy ? entropy[i] : 0u, hc);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
    uint32_t w[8];
    hc = 0x8b51f9ddu;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i & 3] ^ hc;
        hc *= 0x58f38dedu;
        v *= hc;
        v ^= v >> 16;
        w[i] = v;
    }
    uint64_t u0 = (uint64_t)w[0] | ((uint64_t)w[1] << 32), u1 = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
    uint64_t u2 = (uint64_t)w[4] | ((uint64_t)w[5] << 32), u3 = (uint64_t)w[6] | ((uint64_t)w[7] << 32);
    // initstate = u0:u1 (hi:lo), initseq = u2:u3; inc = (initseq << 1) | 1
    r.inc_hi = (u2 << 1) | (u3 >> 63);
    r.inc_lo = (u3 << 1) | 1ULL;
    r.s_lo = 0;
    r.s_hi = 0;
    pcg_step(r.s_lo, r.s_hi, r.inc_lo, r.inc_hi);
    u128_add(r.s_lo, r.s_hi, u1, u0);
    pcg_step(r.s_lo, r.s_hi, r.inc_lo, r.inc_hi);
    r.uinteger = 0;
    r.loaded = true;
    r.dirty = true;
}

// Product-defined counter-based action stream for rollouts (DESIGN.md "synthetic actions");
// splitmix64 finaliser over (seed, env, frame, agent), multiply-shift to [0, n_actions).
__device__ __forceinline__ int synth_action(uint64_t action_seed, uint64_t global_env, uint64_t frame, int agent,
                                            uint32_t n_actions) {
    uint64_t z = action_seed + 0x9E3779B97F4A7C15ULL * (2ULL * global_env + (uint64_t)agent + 1ULL);
    z ^= frame * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int)(((z >> 32) * (uint64_t)n_actions) >> 32);
}

}  // namespace pz
