// Device-side pieces shared by the kernel translation units: launch parameters, observation output
// (shared-memory staging + one bulk async copy per warp), statistics, rewards, and the per-step
// kernel template. The step kernel is instantiated per (AI_MASK, observation dtype) in
// pz_step_ai{0,1,2,3}.cu so that every combination keeps its own register budget (the int32 /
// no-computer instantiation stays at 80 registers) and the translation units compile in parallel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "../../include/pikazoo_b200.h"
#include "pz_physics.cuh"

namespace pz {

constexpr int kThreads = 128;  // 4 warps; 4 x 8960 B of observation staging per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kObsRow = 70;                      // int32 per env: [obs_p1 | obs_p2]
constexpr int kWarpObsBytes = 32 * kObsRow * 4;  // 8960, multiple of 16

struct KParams {
    int32_t *state;
    int64_t n;           // envs in the state buffer (SoA stride)
    int64_t begin, end;  // env range processed by this launch (begin % 32 == 0)
    const void *actions;
    void *obs;
    void *reward;
    uint8_t *done;
    unsigned long long *stats;
    double2 *ep_return;   // RecordEpisodeStatistics running returns (in/out), may be null
    int32_t *ep_length;   // may be null
    uint8_t *truncated;   // may be null
    uint8_t *status;      // may be null: (player_1 base reward + 1) | done << 2 | truncated << 3, one byte per env
    uint32_t *seq;        // may be null (single-CTA launches only): completion word, see pz_episode_io.seq_dev
    uint32_t seq_value;
    StepCfg cfg;
    int autoreset, simplify, shaped, act_dtype, rew_dtype, obs_dtype, normalize;
    int obs_layout, obs_rows;  // PZ_LAYOUT_*; FEATURE_MAJOR: rows per agent (leading dimension = n)
    int max_frames;       // 0: never truncate
    uint64_t state_policy, out_policy;  // L2 cache policies (pz_state.cuh), kL2EvictNormal when hints are off
    int pdl;              // launch the step kernel with programmatic stream serialization
    int x_line, y_line;
    // rollout only
    int K, action_source;
    uint64_t action_seed, first_env, frame0;
    // RewardByBallPosition fused: table[agent][own base reward + 1][zone], evaluated on the host in
    // double exactly as Python evaluates `int + float` (reward_by_ball_position.py:28-29)
    double table[24];
};

// ---- observation output ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Whole warps write their 32 observation rows ([obs_p1 | obs_p2], 70 elements) as ONE contiguous bulk
// copy shared -> global issued by lane 0 (cp.async.bulk; no per-thread strided stores): 8,960 B for
// 4-byte elements, 4,480 B for 2-byte ones. Rows are staged with 64-bit (resp. 32-bit) shared stores;
// the row stride of 70 (resp. 35) words makes them conflict-free per half-warp (resp. warp).
template <int DT>
struct ObsType;
template <>
struct ObsType<PZ_OBS_I32> { static constexpr int bytes = 4; using elem = int; };
template <>
struct ObsType<PZ_OBS_I16> { static constexpr int bytes = 2; using elem = short; };
template <>
struct ObsType<PZ_OBS_F32> { static constexpr int bytes = 4; using elem = float; };
template <>
struct ObsType<PZ_OBS_F16> { static constexpr int bytes = 2; using elem = unsigned short; };
template <>
struct ObsType<PZ_OBS_BF16> { static constexpr int bytes = 2; using elem = unsigned short; };
template <>
struct ObsType<PZ_OBS_F64> { static constexpr int bytes = 8; using elem = double; };

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// Element K of the distinct values as the 16 bits of a float16 / bfloat16. The twelve one-hot and flag elements
// (states, power-hit key, ball power hit: bounds [0, 1], so normalised or not they are 0.0 or 1.0) are a select of
// the constant — no int -> float -> 2-byte conversion chain.
template <int DT, int K>
__device__ __forceinline__ unsigned short obs_bits16(const int (&u)[35], bool normalize) {
    static_assert(DT == PZ_OBS_F16 || DT == PZ_OBS_BF16, "2-byte floating-point rows");
    constexpr bool flag = (K < 26 && (K % 13) >= 7) || K == 34;
    if (flag) return (unsigned short)(u[K] ? (DT == PZ_OBS_F16 ? 0x3C00u : 0x3F80u) : 0u);
    const float f = obs_float<float, K>(u, normalize);
    return DT == PZ_OBS_F16 ? __half_as_ushort(__float2half_rn(f)) : __bfloat16_as_ushort(__float2bfloat16_rn(f));
}

// The 70 elements of one env's row, written with the widest naturally aligned stores
// (row addresses are multiples of 70 * element size, so 8 / 4 / 16-byte aligned).
template <int DT>
__device__ __forceinline__ void write_obs_row(const Env &e, bool normalize, void *row) {
    int u[35];
    obs_values(e, u);
    if (DT == PZ_OBS_I32) {
        int2 *r2 = reinterpret_cast<int2 *>(row);
#pragma unroll
        for (int j = 0; j < 35; j++) r2[j] = make_int2(u[obs_src(2 * j)], u[obs_src(2 * j + 1)]);
    } else if (DT == PZ_OBS_I16) {
        uint32_t *r = reinterpret_cast<uint32_t *>(row);
#pragma unroll
        for (int j = 0; j < 35; j++)
            r[j] = ((uint32_t)u[obs_src(2 * j)] & 0xFFFFu) | ((uint32_t)u[obs_src(2 * j + 1)] << 16);
    } else if (DT == PZ_OBS_F64) {
        double f[35];
        obs_floats(u, f, normalize);
        double2 *r = reinterpret_cast<double2 *>(row);
#pragma unroll
        for (int j = 0; j < 35; j++) r[j] = make_double2(f[obs_src(2 * j)], f[obs_src(2 * j + 1)]);
    } else {
        float f[35];
        obs_floats(u, f, normalize);
        if (DT == PZ_OBS_F32) {
            float2 *r = reinterpret_cast<float2 *>(row);
#pragma unroll
            for (int j = 0; j < 35; j++) r[j] = make_float2(f[obs_src(2 * j)], f[obs_src(2 * j + 1)]);
        } else {
            uint32_t *r = reinterpret_cast<uint32_t *>(row);
#pragma unroll
            for (int j = 0; j < 35; j++)
                r[j] = DT == PZ_OBS_F16 ? pack_half2(f[obs_src(2 * j)], f[obs_src(2 * j + 1)])
                                        : pack_bf162(f[obs_src(2 * j)], f[obs_src(2 * j + 1)]);
        }
    }
}

__device__ __forceinline__ void bulk_store_issue(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_addr(ssrc)), "r"(bytes), "l"(policy)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Returns true if this lane issued a bulk copy it must wait for before the CTA's smem dies.
template <int DT>
__device__ __forceinline__ bool emit_obs_as(const Env &e, bool valid, bool normalize, void *obs, int64_t env_idx,
                                            int64_t end, int *warp_stage, int lane, uint64_t policy) {
    constexpr int row_bytes = kObsRow * ObsType<DT>::bytes;
    const int64_t warp_first = env_idx - lane;
    char *g = reinterpret_cast<char *>(obs);
    if (DT != PZ_OBS_F64 && warp_stage != nullptr && warp_first + 32 <= end) {  // warp-uniform: full warp
        write_obs_row<DT>(e, normalize, reinterpret_cast<char *>(warp_stage) + lane * row_bytes);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            bulk_store_issue(g + warp_first * row_bytes, warp_stage, 32 * row_bytes, policy);
            return true;
        }
    } else if (valid) {  // ragged tail, no staging buffer, float64 rows (17,920 B per warp): plain vector stores
        write_obs_row<DT>(e, normalize, g + env_idx * row_bytes);
    }
    return false;
}

// ENV_MAJOR_SHARED: obs[n][35], player_1's row only (player_2's observation is the same 35 values with the two
// player blocks swapped, pikazoo_env.py:585-586): a quarter of the int32 [n][2][35] bytes as int16. Whole warps
// stage their 32 rows as the exact global image (2,240 or 4,480 contiguous bytes) and lane 0 issues one bulk copy.
template <int DT>
__device__ __forceinline__ void write_obs_row_shared(const Env &e, void *row) {
    static_assert(DT == PZ_OBS_I32 || DT == PZ_OBS_I16, "shared rows are integer");
    int u[35];
    obs_values(e, u);
    if (DT == PZ_OBS_I32) {
        int *r = reinterpret_cast<int *>(row);
#pragma unroll
        for (int j = 0; j < 35; j++) r[j] = u[j];
    } else {  // 70-byte rows: odd rows are only 2-byte aligned
        short *r = reinterpret_cast<short *>(row);
#pragma unroll
        for (int j = 0; j < 35; j++) r[j] = (short)u[j];
    }
}
template <int DT>
__device__ __forceinline__ bool emit_obs_shared_as(const Env &e, bool valid, void *obs, int64_t env_idx, int64_t end,
                                                   int *warp_stage, int lane, uint64_t policy) {
    constexpr int row_bytes = PZ_OBS_WORDS * ObsType<DT>::bytes;
    const int64_t warp_first = env_idx - lane;
    char *g = reinterpret_cast<char *>(obs);
    if (warp_stage != nullptr && warp_first + 32 <= end) {  // warp-uniform: full warp
        write_obs_row_shared<DT>(e, reinterpret_cast<char *>(warp_stage) + lane * row_bytes);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            bulk_store_issue(g + warp_first * row_bytes, warp_stage, 32 * row_bytes, policy);
            return true;
        }
    } else if (valid) {
        write_obs_row_shared<DT>(e, g + env_idx * row_bytes);
    }
    return false;
}

// f(std::integral_constant<int, K>{}) for K = 0, 1, ... (a compile-time loop whose body sees K as a constant)
template <class F, int... K>
__device__ __forceinline__ void for_each_constant(F &&f, std::integer_sequence<int, K...>) {
    (f(std::integral_constant<int, K>{}), ...);
}

// FEATURE_MAJOR: obs[((a * rows) + k) * ld + i]. Consecutive lanes hold consecutive envs, so every one of
// the 70 stores of a warp is one contiguous 64 / 128 / 256-byte segment: no staging, no shared memory.
// Streaming stores (st.global.cs): written once, read by the policy, never by the simulator.
template <int DT>
__device__ __forceinline__ void emit_obs_feature_major(const Env &e, bool valid, bool normalize, void *obs,
                                                       int64_t env_idx, int64_t ld, int rows) {
    if (!valid) return;
    using T = typename ObsType<DT>::elem;
    int u[35];
    obs_values(e, u);
    T *p0 = reinterpret_cast<T *>(obs) + env_idx;  // two running pointers instead of 70 hoisted addresses
    T *p1 = p0 + (int64_t)rows * ld;
    // agent 0's row k is value k; agent 1's rows are [opponent block | own block | ball block]
    auto emit = [&](auto K) {
        constexpr int k = decltype(K)::value;
        T v;
        if constexpr (DT == PZ_OBS_I32 || DT == PZ_OBS_I16)
            v = (T)u[k];
        else if constexpr (DT == PZ_OBS_F64)
            v = (T)obs_float<double, k>(u, normalize);
        else if constexpr (DT == PZ_OBS_F32)
            v = (T)obs_float<float, k>(u, normalize);
        else
            v = (T)obs_bits16<DT, k>(u, normalize);
        constexpr int k1 = k < 13 ? k + 13 : (k < 26 ? k - 13 : k);  // row of value k in agent 1's observation
        __stcs(p0 + (int64_t)k * ld, v);
        __stcs(p1 + (int64_t)k1 * ld, v);
    };
    for_each_constant(emit, std::make_integer_sequence<int, 35>{});
}

// The same rows from a whole CTA of kThreads consecutive envs, transposed through shared memory: every thread
// stages its 35 distinct values as column threadIdx.x of stage[35][kThreads] (35 shared stores at immediate
// offsets instead of 70 global stores with 64-bit address arithmetic each), then every warp writes whole rows —
// kThreads envs of one feature, 256 or 512 contiguous bytes per store instruction — to both agents' copies of
// the row. Needs full CTAs and 4-element-aligned rows (ld % 4 == 0); everything else takes the direct path.
template <int DT>
__device__ __forceinline__ void emit_obs_feature_major_staged(const Env &e, bool normalize, void *obs,
                                                              int64_t cta_first, int64_t ld, int rows, void *stage_raw) {
    using T = typename ObsType<DT>::elem;
    static_assert(sizeof(T) == 2 || sizeof(T) == 4, "staged rows are 2- or 4-byte elements");
    T(*stage)[kThreads] = reinterpret_cast<T(*)[kThreads]>(stage_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int u[35];
    obs_values(e, u);
    auto put = [&](auto K) {
        constexpr int k = decltype(K)::value;
        T v;
        if constexpr (DT == PZ_OBS_I32 || DT == PZ_OBS_I16)
            v = (T)u[k];
        else if constexpr (DT == PZ_OBS_F32)
            v = (T)obs_float<float, k>(u, normalize);
        else
            v = (T)obs_bits16<DT, k>(u, normalize);
        stage[k][tid] = v;
    };
    for_each_constant(put, std::make_integer_sequence<int, 35>{});
    __syncthreads();
    using V = typename std::conditional<sizeof(T) == 2, uint2, uint4>::type;  // four elements per lane
    // Warp w writes rows w, w + 4, ...: row k = w + 4 j goes to agent 0's row k and to agent 1's row k + d(k), d = +13 /
    // -13 / 0 for the own / opponent / ball block. Unrolled over j the offset d is a compile-time constant except in the
    // two iterations that straddle a block boundary (k = 12..15, 24..27), and each address is ONE 32 x 32 + 64-bit
    // multiply-add on a base computed once (the rolled loop with its 64-bit index arithmetic was 15 % of this kernel's
    // instructions, 22 per row).
    const char *srow = reinterpret_cast<const char *>(stage) + (size_t)warp * kThreads * sizeof(T) + lane * sizeof(V);
    const uint32_t ldb = (uint32_t)ld * (uint32_t)sizeof(T);  // bytes per row (the caller checked ld <= 2^28)
    uint64_t g = reinterpret_cast<uint64_t>(obs) + ((int64_t)warp * ld + cta_first + 4 * lane) * (int64_t)sizeof(T);
    uint64_t g1 = g + (uint64_t)(uint32_t)rows * ldb;  // agent 1's copy
    asm volatile("" : "+l"(g), "+l"(g1));  // two materialised bases: every address below is one IMAD.WIDE on one of them
#pragma unroll
    for (int j = 0; j < (35 + kWarps - 1) / kWarps; j++) {
        const int k = warp + kWarps * j;
        if (kWarps * j + kWarps - 1 >= 35 && k >= 35) break;  // only the last iteration is ragged
        const V v = *reinterpret_cast<const V *>(srow + (size_t)j * kWarps * kThreads * sizeof(T));
        int d;  // agent 1's row of value k, minus k
        if (kWarps * j + kWarps - 1 < 13) d = 13;
        else if (kWarps * j >= 13 && kWarps * j + kWarps - 1 < 26) d = -13;
        else if (kWarps * j >= 26) d = 0;
        else d = k < 13 ? 13 : (k < 26 ? -13 : 0);
        __stcs(reinterpret_cast<V *>(g + (uint64_t)(uint32_t)(kWarps * j) * ldb), v);
        __stcs(reinterpret_cast<V *>(g1 + (int64_t)(kWarps * j + d) * (int64_t)ldb), v);
    }
}

__device__ __forceinline__ void emit_obs_feature_major(const Env &e, bool valid, int obs_dtype, bool normalize,
                                                       void *obs, int64_t env_idx, int64_t ld, int rows) {
    switch (obs_dtype) {  // launch-uniform
        case PZ_OBS_I32: emit_obs_feature_major<PZ_OBS_I32>(e, valid, normalize, obs, env_idx, ld, rows); break;
        case PZ_OBS_I16: emit_obs_feature_major<PZ_OBS_I16>(e, valid, normalize, obs, env_idx, ld, rows); break;
        case PZ_OBS_F32: emit_obs_feature_major<PZ_OBS_F32>(e, valid, normalize, obs, env_idx, ld, rows); break;
        case PZ_OBS_F16: emit_obs_feature_major<PZ_OBS_F16>(e, valid, normalize, obs, env_idx, ld, rows); break;
        case PZ_OBS_BF16: emit_obs_feature_major<PZ_OBS_BF16>(e, valid, normalize, obs, env_idx, ld, rows); break;
        default: emit_obs_feature_major<PZ_OBS_F64>(e, valid, normalize, obs, env_idx, ld, rows); break;
    }
}

__device__ __forceinline__ bool emit_obs_shared(const Env &e, bool valid, int obs_dtype, void *obs, int64_t env_idx,
                                                int64_t end, int *warp_stage, int lane, uint64_t policy) {
    if (obs_dtype == PZ_OBS_I32) return emit_obs_shared_as<PZ_OBS_I32>(e, valid, obs, env_idx, end, warp_stage, lane, policy);
    return emit_obs_shared_as<PZ_OBS_I16>(e, valid, obs, env_idx, end, warp_stage, lane, policy);
}

__device__ __forceinline__ bool emit_obs(const Env &e, bool valid, int obs_dtype, bool normalize, void *obs,
                                         int64_t env_idx, int64_t end, int *warp_stage, int lane, uint64_t policy) {
    switch (obs_dtype) {  // launch-uniform
        case PZ_OBS_I32: return emit_obs_as<PZ_OBS_I32>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
        case PZ_OBS_I16: return emit_obs_as<PZ_OBS_I16>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
        case PZ_OBS_F32: return emit_obs_as<PZ_OBS_F32>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
        case PZ_OBS_F16: return emit_obs_as<PZ_OBS_F16>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
        case PZ_OBS_BF16: return emit_obs_as<PZ_OBS_BF16>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
        default: return emit_obs_as<PZ_OBS_F64>(e, valid, normalize, obs, env_idx, end, warp_stage, lane, policy);
    }
}

// ---- statistics ------------------------------------------------------------------------------------
__device__ __forceinline__ void stat_add(unsigned long long *stats, int slot, unsigned v) {
    if (v) atomicAdd(stats + slot, (unsigned long long)v);
}

// Episode-granular events only (rare), aggregated per warp before touching L2 atomics.
__device__ __forceinline__ void accumulate_stats(unsigned long long *stats, const Env &e, bool terminated,
                                                 bool was_reset, bool bad, bool frozen, bool truncated, int lane) {
    if (!__any_sync(kFullMask, terminated || was_reset || bad || frozen || truncated)) return;  // nearly every frame
    const unsigned tm = __ballot_sync(kFullMask, terminated);
    const unsigned rm = __ballot_sync(kFullMask, was_reset);
    const unsigned bm = __ballot_sync(kFullMask, bad);
    const unsigned fm = __ballot_sync(kFullMask, frozen);
    const unsigned xm = __ballot_sync(kFullMask, truncated);
    unsigned frames = 0, s1 = 0, s2 = 0, w1 = 0;
    if (tm) {
        frames = __reduce_add_sync(kFullMask, terminated ? (unsigned)e.ep_frames : 0u);
        s1 = __reduce_add_sync(kFullMask, terminated ? (unsigned)e.score[0] : 0u);
        s2 = __reduce_add_sync(kFullMask, terminated ? (unsigned)e.score[1] : 0u);
        w1 = __popc(__ballot_sync(kFullMask, terminated && e.score[0] > e.score[1]));
    }
    if (lane == 0) {
        stat_add(stats, PZ_STAT_EPISODES, __popc(tm));
        stat_add(stats, PZ_STAT_EPISODE_FRAMES, frames);
        stat_add(stats, PZ_STAT_P1_WINS, w1);
        stat_add(stats, PZ_STAT_P2_WINS, __popc(tm) - w1);
        stat_add(stats, PZ_STAT_P1_POINTS, s1);
        stat_add(stats, PZ_STAT_P2_POINTS, s2);
        stat_add(stats, PZ_STAT_RESETS, __popc(rm));
        stat_add(stats, PZ_STAT_BAD_ACTIONS, __popc(bm));
        stat_add(stats, PZ_STAT_FROZEN, __popc(fm));
        stat_add(stats, PZ_STAT_TRUNCATED, __popc(xm));
    }
}

__device__ __forceinline__ void load_actions(const KParams &P, int64_t i, int &a1, int &a2) {
    if (P.act_dtype == PZ_ACT_I32) {
        int2 a = reinterpret_cast<const int2 *>(P.actions)[i];
        a1 = a.x;
        a2 = a.y;
    } else if (P.act_dtype == PZ_ACT_I64) {
        longlong2 a = reinterpret_cast<const longlong2 *>(P.actions)[i];
        a1 = (a.x < -1 || a.x > 1000) ? -1 : (int)a.x;
        a2 = (a.y < -1 || a.y > 1000) ? -1 : (int)a.y;
    } else {
        uchar2 a = reinterpret_cast<const uchar2 *>(P.actions)[i];
        a1 = a.x;
        a2 = a.y;
    }
}

// RewardByBallPosition zone (reward_by_ball_position.py:22-26) from the post-step ball
__device__ __forceinline__ int ball_zone(const Env &e, const KParams &P) {
    return (e.b.y > P.y_line ? 1 : 0) + 2 * (e.b.x >= P.x_line ? 1 : 0);
}

// The (wrapped) rewards of one executed step: RewardByBallPosition / RewardInNormalState come from
// the host-built table [agent][own base reward + 1][zone].
__device__ __forceinline__ void step_rewards(const KParams &P, const Env &e, int base, double &r1, double &r2) {
    if (P.shaped) {
        const int z = ball_zone(e, P);
        r1 = P.table[(base + 1) * 4 + z];
        r2 = P.table[12 + (1 - base) * 4 + z];
    } else {
        r1 = (double)base;
        r2 = (double)(-base);
    }
}

template <bool F32_ONLY = false>
__device__ __forceinline__ void store_reward(const KParams &P, int64_t i, double r1, double r2) {
    if (F32_ONLY || P.rew_dtype == PZ_REW_F32) {
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(
                         reinterpret_cast<float2 *>(P.reward) + i),
                     "f"((float)r1), "f"((float)r2), "l"(P.out_policy)
                     : "memory");
    } else {
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(
                         reinterpret_cast<double2 *>(P.reward) + i),
                     "d"(r1), "d"(r2), "l"(P.out_policy)
                     : "memory");
    }
}

// An episode is over when the game ended or (optionally) when it reached max_frames step() calls.
__device__ __forceinline__ bool episode_truncated(const KParams &P, const Env &e) {
    return P.max_frames > 0 && !e.game_ended && e.ep_frames >= P.max_frames;
}

// ---- per-step kernel ---------------------------------------------------------------------------
// One thread per env, one tile of 128 envs per CTA, six CTAs resident per SM (shared-memory staging and
// 80 registers both allow exactly six). Measured alternatives that lost (B200, 1,048,576 envs, DESIGN.md
// §4): 2 or 4 software-prefetched tiles per CTA (73 / 75 us against 62 us: the extra live registers or
// spills cost more than the hidden latency gains; the hardware's CTA turnover already overlaps loads
// with the other CTAs' arithmetic), and an evict-last L2 policy on the state words (+2 us).
constexpr int kAiScratchInts = 320;  // computer_decide: 32 x int4 inputs + 32 x 6 results per warp

// Resident CTAs per SM the register allocation is held to (0 = leave it to ptxas). Without computer
// players: 2-byte observation rows stage half as much shared memory, so seven CTAs fit if the registers
// do (72 with a few spills to L1; the kernel waits on its loads, so warps in flight count); feature-major rows stage
// 35 x 128 elements per CTA and run best at seven as well. Measured sweeps in DESIGN.md §4.
#ifndef PZ_FM_MIN_CTAS
#define PZ_FM_MIN_CTAS 7  // bf16 rows per million envs: 52.4 us (6: 58.0, 8: 55.5; direct stores instead of the staged rows: 54.7)
#endif
#ifndef PZ_AI_MIN_CTAS
#define PZ_AI_MIN_CTAS 5  // 94 registers, no spills: 72 us per million envs (4: 101 registers, 79 us; 6: 80 with spills, 77 us)
#endif
#ifndef PZ_HALF_MIN_CTAS
#define PZ_HALF_MIN_CTAS 7  // 72 registers, 28 B of spill loads: 47.6 us per million envs (6: 80 registers, 49.6 us; 8: 64, 53.4 us)
#endif
template <int AI_MASK, int OBS_DT, int LAYOUT, bool PLAIN>
constexpr int step_min_ctas() {
    if (AI_MASK != 0) return PZ_AI_MIN_CTAS;  // (PLAIN at six CTAs / 80 registers: 81.5 against 74.9 us)
    // the PLAIN instantiations need fewer registers: with 2-byte elements eight CTAs of 64 registers now beat seven of 72
    // (fp16 rows 39.7 -> 38.9 us, bf16 feature-major 43.6 -> 42.8 us per million envs; the general ones lose at eight)
    constexpr int two_byte = PLAIN && ObsType<OBS_DT>::bytes == 2 ? 1 : 0;
    if (LAYOUT == PZ_LAYOUT_FEATURE_MAJOR) return OBS_DT == PZ_OBS_F64 ? 4 : PZ_FM_MIN_CTAS + two_byte;
    if (LAYOUT == PZ_LAYOUT_ENV_MAJOR_SHARED) return PZ_HALF_MIN_CTAS;
    return ObsType<OBS_DT>::bytes == 2 ? PZ_HALF_MIN_CTAS + two_byte : (OBS_DT == PZ_OBS_F64 ? 4 : 6);
}

// PLAIN: the configuration of a plain batched run — observations, float32 rewards, done flags and statistics all
// written, auto-reset on, the full action set, and none of the options: no episode returns / lengths, truncated flags,
// status byte or completion word, no frame cap, no shaped rewards, no SimplifyAction. The launch-uniform tests of all of
// these are compiled out (launch_dt in pz_step_inst.inc decides). Worth
// 1 % (int32 rows) to 10 % (bf16 feature-major rows): the options cost registers more than instructions.
template <int AI_MASK, int OBS_DT, int LAYOUT, bool PLAIN = false>
__global__ void __launch_bounds__(kThreads, step_min_ctas<AI_MASK, OBS_DT, LAYOUT, PLAIN>())
    pz_step_kernel(const __grid_constant__ KParams P) {
    // ENV_MAJOR stages the observation rows here (2-byte elements need half the room); FEATURE_MAJOR only
    // needs the computer players' scratch
    constexpr int kRowElems = LAYOUT == PZ_LAYOUT_ENV_MAJOR_SHARED ? PZ_OBS_WORDS : kObsRow;
    constexpr int kWarpRowInts = OBS_DT == PZ_OBS_F64 ? 0 : 32 * kRowElems * ObsType<OBS_DT>::bytes / 4;  // f64 rows are not staged
    constexpr int kStageInts = LAYOUT != PZ_LAYOUT_FEATURE_MAJOR
                                   ? (kWarpRowInts > kAiScratchInts ? kWarpRowInts : kAiScratchInts)
                                   : (AI_MASK != 0 ? kAiScratchInts : 4);
    __shared__ __align__(128) int stage[kWarps][kStageInts];
#ifdef PZ_FM_NO_STAGING
    constexpr bool kFmStaged = false;
#else
    constexpr bool kFmStaged = LAYOUT == PZ_LAYOUT_FEATURE_MAJOR && OBS_DT != PZ_OBS_F64;
#endif
    __shared__ __align__(16) unsigned char fm_stage[kFmStaged ? 35 * kThreads * ObsType<OBS_DT>::bytes : 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = P.begin + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = i < P.end;
    // PLAIN float rows are normalised rows (launch_dt): the raw-float form of the row code is not even compiled
    const bool normalize = PLAIN ? (OBS_DT != PZ_OBS_I32 && OBS_DT != PZ_OBS_I16) : (P.normalize != 0);
    // Programmatic dependent launch: let the next launch in the stream be scheduled into the SM slots
    // this grid's tail frees (it parks in its own cudaGridDependencySynchronize), and wait here for the
    // previous launch to have completed and flushed before the first read. Both are no-ops when the
    // launch did not ask for programmatic stream serialization.
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();

    DrawCtxT<AI_MASK == 0> d;  // computer players: the stream is loaded up front
    d.s = state_ptrs(P.state, P.n, P.state_policy);
    d.idx = i;
    d.r.loaded = false;
    d.r.dirty = false;
    Env e;
    int a1 = 0, a2 = 0;
    // Lanes past the end of a ragged last warp read the last env of the range (a duplicate load) and do nothing with
    // it — every effect below is gated by `valid`. (Giving them a fresh_env instead cost every warp the merge of its 45
    // fields: 3 % of the kernel's instructions.)
    const int64_t il = valid ? i : P.end - 1;
    load_env(e, d.s, il);
    if (AI_MASK == 0 || P.actions) load_actions(P, il, a1, a2);  // (launch_step: only two computer players may omit them)
    if (AI_MASK != 0) rng_load(d.r, d.s, il);  // computer players draw on most frames

    const bool over = e.game_ended || (!PLAIN && episode_truncated(P, e));
    const bool run = valid && !over;
    const bool do_reset = valid && over && (PLAIN || P.autoreset);
    const bool frozen = !PLAIN && valid && over && !P.autoreset;
    const unsigned mask = __ballot_sync(kFullMask, run);
    int base = 0;
    bool bad = false;
    if (run) {
        bool bad1, bad2;
        Input in1, in2;
        if (!PLAIN && P.simplify) {
            in1 = decode_input<0, true>(a1, e.p[0], bad1);
            in2 = decode_input<1, true>(a2, e.p[1], bad2);
        } else {
            in1 = decode_input<0, false>(a1, e.p[0], bad1);
            in2 = decode_input<1, false>(a2, e.p[1], bad2);
        }
        bad = bad1 || bad2;
        // the sprite animation from the compile-time table in global memory (pz_physics.cuh:g_anim_table, read through
        // L1) where that measured faster: 2-byte rows 46.4 -> 45.0 us, computer players 79.2 -> 78.4, int32 rows
        // 59.8 -> 59.5 us per million envs — but float32 rows 66.5 -> 67.7 us, which keep the arithmetic
        constexpr bool kAnimLut = PZ_STEP_ANIM_LUT && OBS_DT != PZ_OBS_F32 && OBS_DT != PZ_OBS_F64;
        // (PLAIN with computer players implies the memoised tables: launch_dt)
        base = step_frame_inputs<AI_MASK, DrawCtxT<AI_MASK == 0>, false, kAnimLut, PLAIN && AI_MASK != 0>(
            mask, e, d, P.cfg, in1, in2, stage[warp], g_anim_table.v);
    } else if (do_reset) {
        reset_env(e, d, P.cfg);
    }

    __syncwarp();  // the staging buffer doubled as the computer players' scratch
    bool pending = false;
    if (PLAIN || P.obs) {
        if (LAYOUT == PZ_LAYOUT_ENV_MAJOR)
            pending = emit_obs_as<OBS_DT>(e, valid, normalize, P.obs, i, P.end, stage[warp], lane, P.out_policy);
        else if constexpr (LAYOUT == PZ_LAYOUT_ENV_MAJOR_SHARED)
            pending = emit_obs_shared_as<OBS_DT>(e, valid, P.obs, i, P.end, stage[warp], lane, P.out_policy);
        else {
            bool staged = false;
            if constexpr (kFmStaged) {
                // CTA-uniform: a full tile, 4-element-aligned rows, row pitch in bytes below 2^32
                if ((i - threadIdx.x) + kThreads <= P.end && (P.n & 3) == 0 && P.n <= (int64_t(1) << 28)) {
                    emit_obs_feature_major_staged<OBS_DT>(e, normalize, P.obs, i - threadIdx.x, P.n, P.obs_rows, fm_stage);
                    staged = true;
                }
            }
            if (!staged) emit_obs_feature_major<OBS_DT>(e, valid, normalize, P.obs, i, P.n, P.obs_rows);
        }
    }
    const bool truncated = !PLAIN && valid && episode_truncated(P, e);  // this call's step reached the cap, or frozen there
    if (valid) {
        if (run || do_reset) {
            store_env(e, d.s, i);
            if (d.r.dirty) rng_store(d.r, d.s, i);
        }
        double r1 = 0.0, r2 = 0.0;
        if (run) {
            if (PLAIN)
                r1 = (double)base, r2 = (double)(-base);
            else
                step_rewards(P, e, base, r1, r2);
        }
        if (PLAIN || P.reward) store_reward<PLAIN>(P, i, r1, r2);
        if (!PLAIN && P.ep_return) {  // record_episode_statistics.py:24-25 (reset zeroes), :32 (step adds)
            if (do_reset) {
                P.ep_return[i] = make_double2(0.0, 0.0);
            } else if (run) {
                double2 acc = P.ep_return[i];
                acc.x += r1;
                acc.y += r2;
                P.ep_return[i] = acc;
            }
        }
        if (!PLAIN && P.ep_length) P.ep_length[i] = e.ep_frames;
        if (PLAIN || P.done) P.done[i] = (uint8_t)(e.game_ended ? 1 : 0);  // a reset cleared it; frozen envs keep it
        if (!PLAIN && P.truncated) P.truncated[i] = (uint8_t)(truncated ? 1 : 0);
        if (!PLAIN && P.status) P.status[i] = (uint8_t)((run ? base + 1 : 1) | (e.game_ended ? 4 : 0) | (truncated ? 8 : 0));
    }
    if (PLAIN || P.stats) {
        accumulate_stats(P.stats, e, run && e.game_ended, do_reset, bad, frozen, run && truncated, lane);
        if (i == P.begin) atomicAdd(P.stats + PZ_STAT_CALLS, (unsigned long long)(P.end - P.begin));
    }
    if (pending) bulk_store_wait_read();
    if (!PLAIN && P.seq != nullptr) {  // launch-uniform; one CTA (checked on the host)
        if (pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the rows have left, not only been read
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            *reinterpret_cast<volatile uint32_t *>(P.seq) = P.seq_value;
        }
    }
}

// Launches pz_step_kernel<AI_MASK, P.obs_dtype, P.obs_layout> over n envs; defined in pz_step_ai*.cu.
template <int AI_MASK>
void launch_step_kernel(int64_t n_envs, cudaStream_t st, const KParams &P);

}  // namespace pz
