// pz_rollout_policy: K frames of `observation -> MLP policy -> sampled actions -> raw_env.step` per launch with the
// env, its PCG64 stream, the observation tile, the hidden activations and the logits all on chip: BASELINE.json
// configs[4] ("actions from an on-device MLP policy in a rollout loop") without the per-frame round trip of the
// observations through HBM that the two-kernel loop (pz_step + pz_policy_mlp_act) makes — 140 B written and 160 B
// read back per env and frame there, 0 here; HBM sees the packed state once per K frames (plus, optionally, the two
// sampled action bytes per frame for a trajectory buffer).
//
// One persistent CTA per SM; kGroups groups of 128 threads; thread t of a group IS env t of the group's current
// 128-env tile for all K frames (state in registers), for both agents. Per frame and tile:
//
//   obs       the thread writes its env's 35 features + the bias input as ONE row of the K-major A operand in shared
//             memory (bf16, 5 x 16-byte stores at thread * 16: conflict-free, no transposition). NormalizeObservation
//             is fused as table look-ups: every bounded field indexes a bf16 table built at kernel start with the very
//             arithmetic of the step kernel's observation output (bit-identical rows). player_2's observation is a
//             permutation of player_1's (pikazoo_env.py:585-586), so instead of a second tile the COLUMNS of player_2's
//             W1 are permuted once, at staging: both agents' first layers are one MMA batch on one tile.
//   layer 1   D1[128 envs][160] = X[128][48] . [W1_p1 ; W1_p2']^T   3 tcgen05.mma (M 128, N 160, K 16), fp32 in TMEM
//   epilogue  thread t reads row t of D1 (tcgen05.ld), rectifies, rounds to bf16 and stores the packed pairs back
//             over the columns it has read (tcgen05.st): H never leaves TMEM
//   layer 2   D2_a[128][32] = H_a[128][80] . W2_a^T, A from TMEM     2 x 5 tcgen05.mma
//   sample    thread t reads its env's two rows of logits and inverts the cumulative distribution in registers
//             (sample_inverse_cdf, pz_policy.cuh — the sampler and counter stream of pz_policy_mlp_act)
//   step      decode_keys + step_frame (pz_physics.cuh), auto-reset as pz_rollout
//
// A frame of one tile is a serial chain (obs -> MMA -> epilogue -> MMA -> sample -> step); the SM overlaps the
// chains of its kGroups tiles. Only the stretch from layer 1 to the logits needs tensor memory (160 columns), so
// the groups share kSlots = 3 column slots (480 of the SM's 512 columns) through a free-mask in shared memory:
// a group takes a slot before it issues layer 1 and returns it as soon as its threads hold their logits.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#include "../../include/pikazoo_b200.h"
#include "pz_kernels.cuh"
#include "pz_physics.cuh"
#include "pz_tcgen05.cuh"

namespace pzrp {

using namespace pz;
using pzp::tc::elect_one;
using pzp::tc::instr_desc;
using pzp::tc::mbar_wait;
using pzp::tc::mma_commit;
using pzp::tc::mma_ss;
using pzp::tc::mma_ts;
using pzp::tc::relu_pack;
using pzp::tc::smem_desc;
using pzp::tc::smem_u32;
using pzp::tc::tc_fence_after;
using pzp::tc::tc_fence_before;
using pzp::tc::tmem_ld16;
using pzp::tc::tmem_ld2;
using pzp::tc::tmem_ld8;
using pzp::tc::tmem_ld_wait;
using pzp::tc::tmem_st8;
using pzp::tc::tmem_st_wait;

// Tiles in flight per SM. Measured on B200, configs[4] (2 M envs, K = 64): 3 groups 18.1 G env-steps/s (every group its
// own slot, 170 registers), 4: 20.8, 5: 21.9 (96 registers), 6: 20.4 (80 registers, 420 B of spills). With the PCG64
// stream and the statistics parked in shared memory (DrawCtxParked): 5: 21.5, 6: 23.1 (150 B of spills), 7: 23.3 —
// six it is; parking five more cold env fields (boldness, stand-by, landing point) measured slower (21.9 at six).
// Again after the select-form player_move, the animation table and the PLAIN instantiation: 5: 24.3, 6: 26.05, 7: 25.9.
// The profile (profiles/r02_ncu_full_rollout_policy.txt): 1,020 warp instructions per warp and frame, issue slots 67 %
// busy with five warps per scheduler of which one or two sit in a barrier, a slot wait or an MMA wait at any time:
// bound by how many warps can run, i.e. by registers. Tried and dropped: staggering the two agents' chains (layer 2 of
// player_1 during player_2's epilogue, layer 2 of player_2 during player_1's sampling: one more barrier, the slot held
// through a sampler, and the two samplers no longer interleave: 20.2 G); four slots of 128 columns (hidden padded to 64,
// timing only: +3 %).
#ifndef PZ_RP_GROUPS
#define PZ_RP_GROUPS 6
#endif
#ifndef PZ_RP_BACKOFF_NS
#define PZ_RP_BACKOFF_NS 100  // sleep between polls of the MMA barrier: the kernel is bound by instruction issue
#endif
constexpr int kGroups = PZ_RP_GROUPS;  // tiles in flight per SM
constexpr int kGroupThreads = 128;
constexpr int kThreads = kGroups * kGroupThreads;
constexpr int kTileEnvs = 128;
constexpr int kKP = PZ_POLICY_MAX_FEATURES;  // 48 = 3 k-steps of 16
#ifndef PZ_RP_HIDDEN_PAD
#define PZ_RP_HIDDEN_PAD PZ_POLICY_MAX_HIDDEN
#endif
constexpr int kHP = PZ_RP_HIDDEN_PAD;        // 80 per agent; N of layer 1 = 160, 5 k-steps of layer 2
constexpr int kAP = 32;                      // N of layer 2 per agent
constexpr int kSlots = 512 / (2 * kHP);  // 3
constexpr uint32_t kSlotCols = 2 * kHP;  // D1 0..159 | H 0..79 (in place) | D2 of agent a at 80 + 32 a
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD2 = kHP;
static_assert(kSlots * kSlotCols <= kTmemCols && kHP + 2 * kAP <= 2 * kHP, "TMEM columns");
static_assert(2 * kGroups + 1 <= 16, "named barriers");  // (a timing experiment with 4 slots of 128 columns gained 3 %)

// shared memory (bytes). Canonical no-swizzle K-major layouts, 8 x 16-byte core matrices:
//   X  (A): element (env m, feature k)       at (k / 8) * 2048 + m * 16 + (k % 8) * 2
//   W1 (B): element (hidden row n, feature k) at (k / 8) * 2560 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2, n = 80 a + unit
//   W2 (B): element (action n, hidden k)     at (k / 8) *  512 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2
constexpr int kXKGroup = kTileEnvs * 16, kXTile = (kKP / 8) * kXKGroup;     // 2048, 12288
constexpr int kW1KGroup = (2 * kHP / 8) * 128, kW1Bytes = (kKP / 8) * kW1KGroup;  // 2560, 15360
constexpr int kW2KGroup = (kAP / 8) * 128, kW2Agent = (kHP / 8) * kW2KGroup;      // 512, 5120
// observation tables (bf16 bits, uint16 entries)
constexpr int kLutPX = 0, kLutPY = kLutPX + 512, kLutPYV = kLutPY + 256, kLutDive = kLutPYV + 64, kLutLy = kLutDive + 4,
              kLutQ4 = kLutLy + 8, kLutBX = kLutQ4 + 8, kLutBY = kLutBX + 512, kLutX432 = kLutBY + 1024,
              kLutBXV = kLutX432 + 512, kLutEntries = kLutBXV + 64;
constexpr int kOffW1 = 0, kOffW2 = kOffW1 + kW1Bytes, kOffX = kOffW2 + 2 * kW2Agent, kOffLut = kOffX + kGroups * kXTile,
              kOffBar = (kOffLut + 2 * kLutEntries + 15) / 16 * 16, kOffSlot = kOffBar + 8 * kGroups,
              kOffMask = kOffSlot + 4 * kGroups, kOffTmem = kOffMask + 4, kOffAct = kOffTmem + 4,  // [2][32] uint32
              kOffStats = (kOffAct + 2 * 32 * 4 + 7) / 8 * 8,                // uint64 [PZ_NUM_STATS]
              kOffRng = kOffStats + 8 * PZ_NUM_STATS;                       // [group][9 words][128 threads] uint32
constexpr int kRngWords = 9;
constexpr int kOffAnim = kOffRng + kGroups * kRngWords * kGroupThreads * 4;  // uint32 [kAnimLutEntries]: player_animate
constexpr size_t kSmemBytes = kOffAnim + (size_t)pz::kAnimLutEntries * 4;

struct Params {
    int32_t *state;
    int64_t n;
    StepCfg cfg;
    int K, simplify, normalize, max_frames, greedy;
    const __nv_bfloat16 *w1, *w2;  // [2][h1][k1], [2][n_actions][k2]
    int h1, k1, n_actions, k2;
    uint64_t seed, step0, first_env;
    unsigned char *actions_out;  // optional [K][n][2]
    float *logits_out;           // optional [K][n][2][n_actions]
    unsigned long long *stats;
    uint64_t state_policy;
};

__device__ __forceinline__ void group_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kGroupThreads) : "memory");
}
__device__ __forceinline__ void group_arrive(int id) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(kGroupThreads) : "memory");
}

// The env's PCG64 stream parked in shared memory for the K frames: without computer players a frame draws only at a
// round start and on the rare x_velocity == 0 collision, so the stream's nine words need not occupy registers across
// the frame loop — registers are what bounds this kernel (tiles in flight per SM). Word w of thread t lies at
// [w][t]: conflict-free.
struct DrawCtxParked {
    uint32_t *slot;  // this thread's column of [9][128]
    __device__ __forceinline__ void fetch(Rng &r) const {
        r.s_lo = (uint64_t)slot[0] | ((uint64_t)slot[kGroupThreads] << 32);
        r.s_hi = (uint64_t)slot[2 * kGroupThreads] | ((uint64_t)slot[3 * kGroupThreads] << 32);
        r.inc_lo = (uint64_t)slot[4 * kGroupThreads] | ((uint64_t)slot[5 * kGroupThreads] << 32);
        r.inc_hi = (uint64_t)slot[6 * kGroupThreads] | ((uint64_t)slot[7 * kGroupThreads] << 32);
        r.uinteger = slot[8 * kGroupThreads];
        r.loaded = true;
        r.dirty = false;
    }
    __device__ __forceinline__ void park(const Rng &r) const {  // (the increment never changes)
        slot[0] = (uint32_t)r.s_lo, slot[kGroupThreads] = (uint32_t)(r.s_lo >> 32);
        slot[2 * kGroupThreads] = (uint32_t)r.s_hi, slot[3 * kGroupThreads] = (uint32_t)(r.s_hi >> 32);
        slot[8 * kGroupThreads] = r.uinteger;
    }
    __device__ __forceinline__ void park_all(const Rng &r) const {
        park(r);
        slot[4 * kGroupThreads] = (uint32_t)r.inc_lo, slot[5 * kGroupThreads] = (uint32_t)(r.inc_lo >> 32);
        slot[6 * kGroupThreads] = (uint32_t)r.inc_hi, slot[7 * kGroupThreads] = (uint32_t)(r.inc_hi >> 32);
    }
    template <uint32_t HIGH>
    __device__ __forceinline__ int integers(int &has32) {
        Rng r;
        fetch(r);
        const int v = rng_integers<HIGH>(r, has32);
        park(r);
        return v;
    }
    __device__ __forceinline__ void computer_draws(int &has32, bool near, bool search, int &standby, int &y_first) {
        Rng r;
        fetch(r);
        rng_computer_draws(r, has32, near, search, standby, y_first);
        park(r);
    }
    __device__ __forceinline__ void integers5_twice(int &has32, int &a, int &b) {
        Rng r;
        fetch(r);
        rng_integers5_twice(r, has32, a, b);
        park(r);
    }
};

// bf16 bits of feature K for raw value v, by the arithmetic of the step kernel's observation output
// (obs_float<float, K> then round-to-nearest-even): the tables and the one computed feature share it.
template <int K>
__device__ __forceinline__ uint16_t feature_bits(int v, bool normalize) {
    const float f = normalize ? div_const_f32<obs_high(K) - obs_low(K)>(v - obs_low(K)) : (float)v;
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
template <int K>
__device__ __forceinline__ void fill_table(uint16_t *t, int count, int offset, bool normalize) {
    for (int i = threadIdx.x; i < count; i += kThreads) t[i] = feature_bits<K>(i - offset, normalize);
}

// The 36 bf16 features of one env (35 observation values in player_1's order + the bias input 1.0) as 18 packed words
__device__ __forceinline__ void pack_features(const Env &e, const uint16_t *lut, bool normalize, uint32_t (&w)[18]) {
    constexpr uint32_t kOne = 0x3F80u;
    uint32_t h[36];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const Player &p = e.p[i];
        uint32_t *o = h + 13 * i;
        o[0] = lut[kLutPX + p.x];
        o[1] = lut[kLutPY + p.y];
        o[2] = lut[kLutPYV + p.yv + 32];
        o[3] = lut[kLutDive + p.dive + 1];
        o[4] = lut[kLutLy + p.lying + 4];
        o[5] = lut[kLutQ4 + p.frame];
        o[6] = lut[kLutQ4 + p.delay];
#pragma unroll
        for (int s = 0; s < 5; s++) o[7 + s] = p.state == s ? kOne : 0u;
        o[12] = p.keyprev ? kOne : 0u;
    }
    const Ball &b = e.b;
    h[26] = lut[kLutBX + b.x];
    h[27] = lut[kLutBY + b.y + 512];
    h[28] = lut[kLutX432 + b.px];
    h[29] = lut[kLutBY + b.py + 512];
    h[30] = lut[kLutX432 + b.ppx];
    h[31] = lut[kLutBY + b.ppy + 512];
    h[32] = lut[kLutBXV + b.xv + 32];
    h[33] = feature_bits<33>(b.yv, normalize);  // 16-bit range: computed
    h[34] = b.pow ? kOne : 0u;
    h[35] = kOne;  // the bias input (policy.py MLPPolicy.ONES_ROW)
#pragma unroll
    for (int j = 0; j < 18; j++) w[j] = h[2 * j] | (h[2 * j + 1] << 16);
}

// PLAIN: sampled actions, no frame cap, no logits exported (the training-loop configuration; the sampled actions may be
// exported): the launch-uniform options are compiled out of the frame loop. 24.6 -> 25.6 G env-steps/s with the action
// export on configs[4], 24.7 -> 25.8 G without.
template <int NA, bool PLAIN>
__global__ void __launch_bounds__(kThreads, 1) pz_rollout_policy_kernel(const __grid_constant__ Params P) {
    const bool opt_greedy = !PLAIN && P.greedy;
    float *const opt_logits_out = PLAIN ? nullptr : P.logits_out;
    unsigned char *const opt_actions_out = P.actions_out;  // (kept at run time: a store per frame, no measurable cost)
    const int opt_max_frames = PLAIN ? 0 : P.max_frames;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // warp-uniform by construction
    const int grp = warp >> 2, wic = warp & 3, ctid = tid - grp * kGroupThreads;
    const uint32_t bar = s_base + kOffBar + 8 * grp;
    volatile uint32_t *slot_of = reinterpret_cast<volatile uint32_t *>(smem + kOffSlot);
    unsigned *slot_mask = reinterpret_cast<unsigned *>(smem + kOffMask);
    uint16_t *lut = reinterpret_cast<uint16_t *>(smem + kOffLut);
    const int id_sync = 1 + 2 * grp, id_free = 2 + 2 * grp;

    // ---- once per CTA: TMEM, barriers, zeroed operands, observation tables, weights in canonical order ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kOffTmem),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (ctid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
    unsigned long long *s_stats = reinterpret_cast<unsigned long long *>(smem + kOffStats);
    if (tid < PZ_NUM_STATS) s_stats[tid] = 0ULL;
    if (tid == 0) {
        *slot_mask = (1u << kSlots) - 1u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kOffLut / 16; i += kThreads) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    {
        const bool nz = P.normalize != 0;
        fill_table<0>(lut + kLutPX, 512, 0, nz);
        fill_table<1>(lut + kLutPY, 256, 0, nz);
        fill_table<2>(lut + kLutPYV, 64, 32, nz);
        fill_table<3>(lut + kLutDive, 4, 1, nz);
        fill_table<4>(lut + kLutLy, 8, 4, nz);
        fill_table<5>(lut + kLutQ4, 8, 0, nz);
        fill_table<26>(lut + kLutBX, 512, 0, nz);
        fill_table<27>(lut + kLutBY, 1024, 512, nz);
        fill_table<28>(lut + kLutX432, 512, 0, nz);
        fill_table<32>(lut + kLutBXV, 64, 32, nz);
    }
    pz::anim_fill(reinterpret_cast<uint32_t *>(smem + kOffAnim), tid, kThreads);  // sprite animation as a table
    if (tid < 64) {  // action -> pre-decoded input of player tid / 32 (action_key_map, SimplifyAction folded in)
        const int a = tid & 31;
        bool bad;
        uint32_t keys;
        if (NA == 13)
            keys = tid < 32 ? decode_keys<0, true>(a, bad) : decode_keys<1, true>(a, bad);
        else
            keys = tid < 32 ? decode_keys<0, false>(a, bad) : decode_keys<1, false>(a, bad);
        reinterpret_cast<uint32_t *>(smem + kOffAct)[tid] = pack_input(keys);
    }
    __syncthreads();
    {
        // W1: both agents stacked along N; player_2's columns permuted into player_1's feature order
        // (obs_p2[k] = obs_p1[k +- 13] for the two player blocks, pikazoo_env.py:585-586)
        __nv_bfloat16 *w1s = reinterpret_cast<__nv_bfloat16 *>(smem + kOffW1);
        for (int i = tid; i < 2 * P.h1 * P.k1; i += kThreads) {
            const int a = i / (P.h1 * P.k1), rem = i - a * (P.h1 * P.k1), n = rem / P.k1, k = rem - n * P.k1;
            const int kk = (a == 1 && k < 26) ? (k < 13 ? k + 13 : k - 13) : k, row = a * kHP + n;
            if (n < kHP) w1s[((kk >> 3) * kW1KGroup + (row >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2) >> 1] = P.w1[i];
        }
        __nv_bfloat16 *w2s = reinterpret_cast<__nv_bfloat16 *>(smem + kOffW2);
        for (int i = tid; i < 2 * NA * P.k2; i += kThreads) {
            const int a = i / (NA * P.k2), rem = i - a * (NA * P.k2), n = rem / P.k2, k = rem - n * P.k2;
            if (k < kHP) w2s[(a * kW2Agent + (k >> 3) * kW2KGroup + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = P.w2[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + kOffTmem);
    const uint32_t s_x = s_base + kOffX + grp * kXTile;
    unsigned char *x_row = smem + kOffX + grp * kXTile + ctid * 16;
    constexpr uint32_t kIdesc1 = instr_desc(2 * kHP, false), kIdesc2 = instr_desc(kAP, false);
    const int64_t n_tiles = (P.n + kTileEnvs - 1) / kTileEnvs;
    const bool normalize = P.normalize != 0;
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x + (int64_t)gridDim.x * grp; tile < n_tiles; tile += (int64_t)gridDim.x * kGroups) {
        const int64_t i = tile * kTileEnvs + ctid;
        const bool valid = i < P.n;
        DrawCtxParked d;
        d.slot = reinterpret_cast<uint32_t *>(smem + kOffRng) + grp * kRngWords * kGroupThreads + ctid;
        Env e;
        if (valid) {
            const StatePtrs sp = state_ptrs(P.state, P.n, P.state_policy);
            load_env(e, sp, i);
            Rng r;
            rng_load(r, sp, i);
            d.park_all(r);
        } else {
            fresh_env(e);
            Rng r = {};
            d.park_all(r);
        }
        const uint64_t genv = P.first_env + (uint64_t)i;

#pragma unroll 1
        for (int k = 0; k < P.K; k++) {
            // ---- observation row -> shared memory (K-major A operand)
            {
                uint32_t w[18];
                pack_features(e, lut, normalize, w);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<uint4 *>(x_row + c * kXKGroup) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                *reinterpret_cast<uint4 *>(x_row + 4 * kXKGroup) = make_uint4(w[16], w[17], 0u, 0u);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> visible to the MMA's reads
            if (wic == 0) {  // take a column slot for this frame's policy evaluation
                if (elect_one()) {
                    unsigned bit;
                    for (;;) {
                        const unsigned m = *reinterpret_cast<volatile unsigned *>(slot_mask);
                        if (m != 0u) {
                            bit = m & (0u - m);
                            if (atomicAnd(slot_mask, ~bit) & bit) {
                                __threadfence_block();  // acquire: the previous holder's last TMEM reads are behind us
                                break;
                            }
                        } else {
                            __nanosleep(32);
                        }
                    }
                    slot_of[grp] = (uint32_t)(__ffs((int)bit) - 1);
                }
                __syncwarp();
            }
            tc_fence_before();
            group_sync(id_sync);
            const uint32_t t_slot = tmem + slot_of[grp] * kSlotCols;
            const uint32_t t_row = t_slot + ((uint32_t)(wic * 32) << 16);
            if (wic == 0 && elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kKP / 16; ks++)
                    mma_ss(t_slot, smem_desc(s_x + ks * 2 * kXKGroup, kXKGroup, 128),
                           smem_desc(s_base + kOffW1 + ks * 2 * kW1KGroup, kW1KGroup, 128), kIdesc1, ks > 0);
                mma_commit(bar);
            }
            const uint32_t nbase = opt_greedy ? 0u : pzp::noise_base(P.seed, P.step0 + (uint64_t)k, genv);
            mbar_wait<PZ_RP_BACKOFF_NS>(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- relu, round to bf16, back into TMEM as the A operands of layer 2 (H column j = hidden 2j, 2j+1)
            {
                constexpr int kChunks = 2 * kHP / 16;
                uint32_t v[2][16], h[8];
#pragma unroll
                for (int c = 0; c < kChunks; c += 2) {
                    tmem_ld16(t_row + 16 * c, v[0]);
                    tmem_ld16(t_row + 16 * (c + 1), v[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 2; q++) {
#pragma unroll
                        for (int j = 0; j < 8; j++) h[j] = relu_pack(v[q][2 * j], v[q][2 * j + 1]);
                        tmem_st8(t_row + 8 * (c + q), h);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            group_sync(id_sync);
            if (wic == 0 && elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int j = 0; j < kHP / 16; j++)
                        mma_ts(t_slot + kColD2 + a * kAP, t_slot + a * (kHP / 2) + 8 * j,
                               smem_desc(s_base + kOffW2 + a * kW2Agent + j * 2 * kW2KGroup, kW2KGroup, 128), kIdesc2, j > 0);
                mma_commit(bar);
            }
            mbar_wait<PZ_RP_BACKOFF_NS>(bar, phase);
            phase ^= 1u;
            tc_fence_after();
            // ---- both agents' logits into registers, then the slot goes back to the pool
            float lg[2][NA];
            {
                uint32_t v0[16], v1[16];
                tmem_ld16(t_row + kColD2, v0);
                tmem_ld16(t_row + kColD2 + kAP, v1);
                if (NA > 16) {
                    uint32_t x0[2], x1[2];
                    tmem_ld2(t_row + kColD2 + 16, x0);
                    tmem_ld2(t_row + kColD2 + kAP + 16, x1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 16; j < NA; j++) {
                        lg[0][j] = __uint_as_float(x0[j - 16]);
                        lg[1][j] = __uint_as_float(x1[j - 16]);
                    }
                } else {
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j = 0; j < (NA < 16 ? NA : 16); j++) {
                    lg[0][j] = __uint_as_float(v0[j]);
                    lg[1][j] = __uint_as_float(v1[j]);
                }
            }
            tc_fence_before();
            if (wic == 0) {  // warps 1..3 only signal; warp 0 waits for them and frees the slot
                group_sync(id_free);
                if (elect_one()) {
                    __threadfence_block();  // release
                    atomicOr(slot_mask, 1u << slot_of[grp]);
                }
            } else {
                group_arrive(id_free);
            }
            if (opt_logits_out != nullptr && valid) {
                float *dst = opt_logits_out + (((int64_t)k * P.n + i) * 2) * NA;
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int j = 0; j < NA; j++) dst[a * NA + j] = lg[a][j];
            }
            // ---- the two actions
            int act[2];
#pragma unroll
            for (int a = 0; a < 2; a++) {
                if (opt_greedy) {
                    float best = pzp::pack_key(-INFINITY, 31);
#pragma unroll
                    for (int j = 0; j < NA; j++) best = fmaxf(best, pzp::pack_key(lg[a][j], j));
                    act[a] = 31 - (int)(__float_as_uint(best) & 31u);
                    if (act[a] >= NA) act[a] = 0;
                } else {
                    act[a] = pzp::sample_inverse_cdf<NA>(lg[a], NA, nbase + (uint32_t)(32 * a) * 0x9E3779B9u);
                }
            }
            if (opt_actions_out != nullptr && valid)
                reinterpret_cast<uchar2 *>(opt_actions_out)[(int64_t)k * P.n + i] =
                    make_uchar2((unsigned char)act[0], (unsigned char)act[1]);
            // ---- one call of the env (pz_rollout semantics: auto-reset always on)
            const bool over = e.game_ended || (opt_max_frames > 0 && e.ep_frames >= opt_max_frames);
            if (valid && !over) {
                const uint32_t *acttab = reinterpret_cast<const uint32_t *>(smem + kOffAct);
                const Input in1 = input_from_packed(e.p[0], acttab[act[0]]);
                const Input in2 = input_from_packed(e.p[1], acttab[32 + act[1]]);
                step_frame_inputs<0, DrawCtxParked, true>(0xFFFFFFFFu, e, d, P.cfg, in1, in2, nullptr,
                                                          reinterpret_cast<const uint32_t *>(smem + kOffAnim));
                // episode-granular events (one in hundreds of frames per env): shared-memory atomics, no registers
                if (e.game_ended) {
                    const bool w1 = e.score[0] > e.score[1];
                    atomicAdd(s_stats + PZ_STAT_EPISODES, 1ULL);
                    atomicAdd(s_stats + PZ_STAT_EPISODE_FRAMES, (unsigned long long)e.ep_frames);
                    atomicAdd(s_stats + (w1 ? PZ_STAT_P1_WINS : PZ_STAT_P2_WINS), 1ULL);
                    atomicAdd(s_stats + PZ_STAT_P1_POINTS, (unsigned long long)e.score[0]);
                    atomicAdd(s_stats + PZ_STAT_P2_POINTS, (unsigned long long)e.score[1]);
                } else if (opt_max_frames > 0 && e.ep_frames >= opt_max_frames) {
                    atomicAdd(s_stats + PZ_STAT_TRUNCATED, 1ULL);
                }
            } else if (valid) {
                reset_env(e, d, P.cfg);
                atomicAdd(s_stats + PZ_STAT_RESETS, 1ULL);
            }
        }

        if (valid) {
            const StatePtrs sp = state_ptrs(P.state, P.n, P.state_policy);
            store_env(e, sp, i);
            Rng r;
            d.fetch(r);
            rng_store(r, sp, i);
        }
        const unsigned cnt = __popc(__ballot_sync(0xFFFFFFFFu, valid));
        if (lane == 0) atomicAdd(s_stats + PZ_STAT_CALLS, (unsigned long long)cnt * (unsigned long long)P.K);
    }
    tc_fence_before();
    __syncthreads();
    if (P.stats != nullptr && tid < PZ_NUM_STATS && s_stats[tid] != 0ULL) atomicAdd(P.stats + tid, s_stats[tid]);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

template <int NA, bool PLAIN>
static cudaError_t launch_one(const Params &P, unsigned grid, cudaStream_t stream, int dev) {
    static std::mutex mu;
    static bool attr_set[64] = {};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(pz_rollout_policy_kernel<NA, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kSmemBytes);
            if (e != cudaSuccess) return e;
            attr_set[dev] = true;
        }
    }
    pz_rollout_policy_kernel<NA, PLAIN><<<grid, kThreads, kSmemBytes, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace pzrp

extern "C" int pz_rollout_policy(int32_t *state_dev, int64_t n, const pz_config *cfg, int32_t K, const void *w1_dev,
                                 int32_t hidden_rows, int32_t features, const void *w2_dev, int32_t n_actions,
                                 int32_t w2_cols, uint64_t seed, uint64_t step0, uint64_t first_env, int32_t greedy,
                                 uint8_t *actions_out_dev, float *logits_out_dev, void *obs_dev, int64_t *stats_dev,
                                 void *stream) {
    using namespace pzrp;
    if (!state_dev || !w1_dev || !w2_dev || n < 0 || K < 1) return PZ_E_BADARG;
    if (int rc = pz::check_config(cfg)) return rc;
    if (cfg->is_player1_computer || cfg->is_player2_computer) return PZ_E_BADCONFIG;  // the policy plays both sides
    if (n_actions != (cfg->simplify_action ? 13 : 18)) return PZ_E_BADARG;           // the env's action space
    if (hidden_rows < 1 || hidden_rows > PZ_POLICY_MAX_HIDDEN || features < 36 || features > kKP || w2_cols < 1 || w2_cols > PZ_POLICY_MAX_HIDDEN)
        return PZ_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(state_dev) & 15u) || (reinterpret_cast<uintptr_t>(obs_dev) & 15u)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    Params P;
    memset(&P, 0, sizeof(P));
    P.state = state_dev;
    P.n = n;
    P.cfg.winning_score = cfg->winning_score;
    P.cfg.serve = cfg->serve;
    P.K = K;
    P.simplify = cfg->simplify_action != 0;
    P.normalize = cfg->normalize_observation != 0;
    P.max_frames = cfg->max_episode_frames;
    P.greedy = greedy != 0;
    P.w1 = reinterpret_cast<const __nv_bfloat16 *>(w1_dev);
    P.w2 = reinterpret_cast<const __nv_bfloat16 *>(w2_dev);
    P.h1 = hidden_rows;
    P.k1 = features;
    P.n_actions = n_actions;
    P.k2 = w2_cols;
    P.seed = seed;
    P.step0 = step0;
    P.first_env = first_env;
    P.actions_out = actions_out_dev;
    P.logits_out = logits_out_dev;
    P.stats = reinterpret_cast<unsigned long long *>(stats_dev);
    P.state_policy = pz::kL2EvictNormal;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + kTileEnvs - 1) / kTileEnvs;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    const bool plain = !P.greedy && P.max_frames <= 0 && P.logits_out == nullptr;
    const cudaError_t err = n_actions == 18 ? (plain ? launch_one<18, true>(P, grid, st, dev) : launch_one<18, false>(P, grid, st, dev))
                                            : (plain ? launch_one<13, true>(P, grid, st, dev) : launch_one<13, false>(P, grid, st, dev));
    if (err != cudaSuccess) return (int)err;
    if (obs_dev) return pz::launch_observe(state_dev, n, cfg, obs_dev, st);  // the observation after the last frame
    return 0;
}
