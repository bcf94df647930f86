// sm_100a kernels + C ABI (include/pikazoo_b200.h) of the batched Pikachu-Volleyball simulator.
//
//   pz_step_kernel     (pz_device.cuh, instantiated in pz_step_ai*.cu) one frame per env per launch;
//                      HBM-bound: 2x128-bit state loads, 2x128-bit state stores,
//                      observation rows staged in shared memory and written with one bulk async copy
//                      (TMA engine, cp.async.bulk) per warp, or stored feature-major without staging.
//   pz_rollout_kernel  K frames per launch with the env and its PCG64 stream in registers.
//   pz_reset_kernel / pz_seed_kernel / export / import / the memoised trajectory tables.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/pikazoo_b200.h"
#include "pz_device.cuh"
#include "pz_kernels.cuh"

// Register cap of the rollout kernel. Its shared memory is 1,280 B per warp, so registers alone set the
// occupancy, and the frame loop is latency-bound (issue slots 65 % busy at 20 warps per SM): a tighter cap
// with a few spills to L1 wins. Measured per 64 frames x 1 M envs, computer vs computer (B200):
// 104 regs 2.69 ms | 96 (no spills, 20 warps) 2.36 | 88 2.53 | 80 2.36 | 72 (188 B of spill loads, 28 warps) 2.23 |
// 64 2.23-2.26 | 56 2.43 | 48 3.09. Round 2, steady state, after the PLAIN specialisation and the straight-line
// power-hit look-ups: 64 2.118 | 72 2.040 | 80 2.012 ms with two computer players, while the kernel without
// computer players (synthetic actions) prefers 72 (1.027 against 1.054 ms at 80).
#ifndef PZ_ROLLOUT_MAXNREG
#define PZ_ROLLOUT_MAXNREG 72
#endif
#ifndef PZ_ROLLOUT_ANIM_LUT
#define PZ_ROLLOUT_ANIM_LUT true
#endif
#ifndef PZ_ROLLOUT_MAXNREG_AI
#define PZ_ROLLOUT_MAXNREG_AI 80
#endif

namespace pz {

// Observation rows from non-inlined code, for the kernels that emit them once per launch (reset,
// rollout): keeps the six dtype variants out of their register allocation. The env is re-read from
// the state the calling thread has just stored (same thread, program order), so nothing but a few
// scalars crosses the call.
__device__ __noinline__ bool emit_obs_cold(int32_t *state, int64_t n, int64_t env_idx, int64_t end, int obs_dtype,
                                           bool normalize, void *obs, int *warp_stage, uint64_t state_policy,
                                           uint64_t out_policy, int layout, int rows) {
    const int lane = threadIdx.x & 31;
    const bool valid = env_idx < end;
    Env e;
    if (valid)
        load_env(e, state_ptrs(state, n, state_policy), env_idx);
    else
        fresh_env(e);
    if (layout == PZ_LAYOUT_FEATURE_MAJOR) {
        emit_obs_feature_major(e, valid, obs_dtype, normalize, obs, env_idx, n, rows);
        return false;
    }
    if (layout == PZ_LAYOUT_ENV_MAJOR_SHARED) return emit_obs_shared(e, valid, obs_dtype, obs, env_idx, end, warp_stage, lane, out_policy);
    return emit_obs(e, valid, obs_dtype, normalize, obs, env_idx, end, warp_stage, lane, out_policy);
}

// ---- K-frame register-resident rollout -------------------------------------------------------
// Shared memory holds only the computer players' scratch (1,280 B per warp): the observation rows, written
// once per K frames, go out with plain vector stores, so that registers alone decide the occupancy.
#ifndef PZ_ROLLOUT_THREADS
#define PZ_ROLLOUT_THREADS 128
#endif
constexpr int kRolloutThreads = PZ_ROLLOUT_THREADS;

// PLAIN: no-op actions and no frame cap (configs[3], and every pre-advance) — the frame loop then carries neither the
// action stream and its decode nor the two truncation tests.
// TABLES (with PLAIN and computer players only): the memoised trajectory tables are known to be there.
template <int AI_MASK, bool PLAIN, bool TABLES = false>
__global__ void __maxnreg__(AI_MASK == 3 ? PZ_ROLLOUT_MAXNREG_AI : PZ_ROLLOUT_MAXNREG) pz_rollout_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(16) int stage[kRolloutThreads / 32][kAiScratchInts];
    const int warp = threadIdx.x >> 5;
    const int64_t i = P.begin + (int64_t)blockIdx.x * kRolloutThreads + threadIdx.x;
    const bool valid = i < P.end;

    DrawCtxT<false> d;  // the stream is loaded up front; no pointer stays live across the frame loop
    d.r.loaded = false;
    d.r.dirty = false;
    Env e;
    if (valid) {
        const StatePtrs sp = state_ptrs(P.state, P.n, P.state_policy);
        load_env(e, sp, i);
        rng_load(d.r, sp, i);
    } else {
        fresh_env(e);
    }
    const uint32_t n_actions = P.simplify ? 13u : 18u;
    const uint64_t genv = P.first_env + (uint64_t)i;

    // Episode-granular statistics (an event per env every few hundred frames): shared-memory atomics at the event, one
    // flush per CTA — seven per-thread counters would sit in registers across the frame loop, and registers (the cap
    // above) are what this kernel is short of.
    __shared__ unsigned long long s_stats[PZ_NUM_STATS];
    if (threadIdx.x < PZ_NUM_STATS) s_stats[threadIdx.x] = 0ULL;
    // the players' sprite animation as a table (pz_physics.cuh:player_animate): 4 KB, filled once for K frames
    __shared__ uint32_t s_anim[kAnimLutEntries];
    anim_fill(s_anim, threadIdx.x, kRolloutThreads);
    __syncthreads();

#pragma unroll 1
    for (int k = 0; k < P.K; k++) {
        const bool run = valid && !e.game_ended && (PLAIN || !episode_truncated(P, e));
        const unsigned mask = __ballot_sync(kFullMask, run);
        if (run) {
            Input in1, in2;
            if (PLAIN) {  // action 0 holds no key (pikazoo_env.py:119-141; SimplifyAction maps 0 to 0)
                in1.xdir = in1.ydir = in1.power = 0;
                in2 = in1;
                e.p[0].keyprev = e.p[1].keyprev = 0;
            } else {
                int a1 = 0, a2 = 0;  // PZ_ACTIONS_NOOP
                if (P.action_source == PZ_ACTIONS_SYNTH) {
                    a1 = synth_action(P.action_seed, genv, P.frame0 + (uint64_t)k, 0, n_actions);
                    a2 = synth_action(P.action_seed, genv, P.frame0 + (uint64_t)k, 1, n_actions);
                }
                bool b1, b2;
                if (P.simplify) {
                    in1 = decode_input<0, true>(a1, e.p[0], b1);
                    in2 = decode_input<1, true>(a2, e.p[1], b2);
                } else {
                    in1 = decode_input<0, false>(a1, e.p[0], b1);
                    in2 = decode_input<1, false>(a2, e.p[1], b2);
                }
            }
            step_frame_inputs<AI_MASK, DrawCtxT<false>, PZ_ROLLOUT_ANIM_LUT, PZ_ROLLOUT_ANIM_LUT, TABLES>(mask, e, d, P.cfg, in1, in2, stage[warp], s_anim);
            if (e.game_ended) {
                atomicAdd(s_stats + PZ_STAT_EPISODES, 1ULL);
                atomicAdd(s_stats + PZ_STAT_EPISODE_FRAMES, (unsigned long long)e.ep_frames);
                atomicAdd(s_stats + (e.score[0] > e.score[1] ? PZ_STAT_P1_WINS : PZ_STAT_P2_WINS), 1ULL);
                atomicAdd(s_stats + PZ_STAT_P1_POINTS, (unsigned long long)e.score[0]);
                atomicAdd(s_stats + PZ_STAT_P2_POINTS, (unsigned long long)e.score[1]);
            } else if (!PLAIN && episode_truncated(P, e)) {
                atomicAdd(s_stats + PZ_STAT_TRUNCATED, 1ULL);
            }
        } else if (valid) {
            reset_env(e, d, P.cfg);
            atomicAdd(s_stats + PZ_STAT_RESETS, 1ULL);
        }
    }

    __syncwarp();
    if (valid) {
        const StatePtrs sp = state_ptrs(P.state, P.n, P.state_policy);
        store_env(e, sp, i);
        if (d.r.dirty) rng_store(d.r, sp, i);
    }
    bool pending = false;
    if (P.obs) pending = emit_obs_cold(P.state, P.n, i, P.end, P.obs_dtype, P.normalize, P.obs, nullptr, P.state_policy,
                                       P.out_policy, P.obs_layout, P.obs_rows);
    __syncthreads();
    if (P.stats) {
        if (threadIdx.x < PZ_NUM_STATS && s_stats[threadIdx.x] != 0ULL) atomicAdd(P.stats + threadIdx.x, s_stats[threadIdx.x]);
        if (i == P.begin)
            atomicAdd(P.stats + PZ_STAT_CALLS, (unsigned long long)(P.end - P.begin) * (unsigned long long)P.K);
    }
    if (pending) bulk_store_wait_read();
}

// ---- reset / seed / export / import -----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pz_reset_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) int stage[kWarps][32 * kObsRow];
    const int warp = threadIdx.x >> 5;
    const int64_t i = P.begin + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool valid = i < P.end;
    DrawCtx d;
    d.s = state_ptrs(P.state, P.n, P.state_policy);
    d.idx = i;
    d.r.loaded = false;
    d.r.dirty = false;
    Env e;
    if (valid) {
        load_env(e, d.s, i);
        reset_env(e, d, P.cfg);
        store_env(e, d.s, i);
        if (d.r.dirty) rng_store(d.r, d.s, i);
    } else {
        fresh_env(e);
    }
    bool pending = false;
    if (P.obs) pending = emit_obs_cold(P.state, P.n, i, P.end, P.obs_dtype, P.normalize, P.obs, stage[warp], P.state_policy,
                                       P.out_policy, P.obs_layout, P.obs_rows);
    if (valid && P.ep_return) P.ep_return[i] = make_double2(0.0, 0.0);
    if (valid && P.ep_length) P.ep_length[i] = 0;
    if (pending) bulk_store_wait_read();
}

// raw_env._get_obs (pikazoo_env.py:576-624) of every env, from the packed state as it stands
__global__ void __launch_bounds__(kThreads) pz_observe_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(128) int stage[kWarps][32 * kObsRow];
    const int warp = threadIdx.x >> 5;
    const int64_t i = P.begin + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool pending = emit_obs_cold(P.state, P.n, i, P.end, P.obs_dtype, P.normalize, P.obs, stage[warp], P.state_policy,
                                       P.out_policy, P.obs_layout, P.obs_rows);
    if (pending) bulk_store_wait_read();
}

__global__ void __launch_bounds__(kThreads)
    pz_seed_kernel(int32_t *state, int64_t n, uint64_t base_seed, uint64_t first_env, const uint64_t *seeds) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    StatePtrs s = state_ptrs(state, n);
    Env e;
    fresh_env(e);
    Rng r;
    pcg64_seed(seeds ? seeds[i] : base_seed + first_env + (uint64_t)i, r);
    store_env(e, s, i);
    rng_store(r, s, i);
    int4 ic;
    ic.x = (int)(uint32_t)r.inc_lo;
    ic.y = (int)(uint32_t)(r.inc_lo >> 32);
    ic.z = (int)(uint32_t)r.inc_hi;
    ic.w = (int)(uint32_t)(r.inc_hi >> 32);
    s.g3[i] = ic;
}

// unpacked layout = oracle/pika_oracle.h pk_env (53 words)
__global__ void __launch_bounds__(kThreads) pz_export_kernel(const int32_t *state, int64_t n, int32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    StatePtrs s = state_ptrs(const_cast<int32_t *>(state), n);
    Env e;
    load_env(e, s, i);
    int32_t *o = out + i * PZ_UNPACKED_WORDS;
    env_to_unpacked(e, o);
    const int4 st = s.g2[i], ic = s.g3[i];
    o[42] = st.x, o[43] = st.y, o[44] = st.z, o[45] = st.w;
    o[46] = ic.x, o[47] = ic.y, o[48] = ic.z, o[49] = ic.w;
    o[51] = (int32_t)s.u[i];
}

__global__ void __launch_bounds__(kThreads) pz_import_kernel(int32_t *state, int64_t n, const int32_t *in) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    StatePtrs s = state_ptrs(state, n);
    const int32_t *o = in + i * PZ_UNPACKED_WORDS;
    Env e;
    env_from_unpacked(e, o);
    store_env(e, s, i);
    s.g2[i] = make_int4(o[42], o[43], o[44], o[45]);
    s.g3[i] = make_int4(o[46], o[47], o[48], o[49]);
    s.u[i] = (uint32_t)o[51];
}

// ---- measurement aid: write-only HBM probe ---------------------------------------------------------
// The per-step kernel is ~87 % stores: the read+write copy bandwidth in MEASURED_PEAKS.json is
// not the ceiling such a kernel sees. Modes (pz_probe_write `mode`):
//   0  128-bit stores, default cache policy, four per thread, one pass over the buffer (a fill kernel)
//   1  the same with the evict-first L2 policy and L1::no_allocate of the step kernel's reward / state stores
//   2  st.global.cs (streaming)
//   3  the step kernel's observation path: every warp fills 8,960 B of shared memory and lane 0 issues one
//      cp.async.bulk shared -> global with the evict-first policy (128-thread CTAs, 35 KB of staging each)
//   4  as 3 without the policy hint
__global__ void __launch_bounds__(256) pz_probe_write_kernel(uint4 *dst, size_t n16, int mode, uint32_t v) {
    const size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t i = base + (size_t)k * 256;
        if (i >= n16) break;
        if (mode == 1)
            asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(dst + i), "r"(v),
                         "l"(kL2EvictFirst)
                         : "memory");
        else if (mode == 2)
            __stcs(dst + i, make_uint4(v, v, v, v));
        else
            dst[i] = make_uint4(v, v, v, v);
    }
}

__global__ void __launch_bounds__(kThreads, 6) pz_probe_bulk_kernel(char *dst, size_t n_blocks, int hint, uint32_t v) {
    __shared__ __align__(128) int stage[kWarps][32 * kObsRow];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t blk = (size_t)blockIdx.x * kWarps + warp;  // one 8,960-byte block per warp
    if (blk >= n_blocks) return;
    int2 *row = reinterpret_cast<int2 *>(stage[warp]) + lane * (kObsRow / 2);
#pragma unroll
    for (int j = 0; j < kObsRow / 2; j++) row[j] = make_int2((int)v + j, lane);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        if (hint)
            bulk_store_issue(dst + blk * kWarpObsBytes, stage[warp], kWarpObsBytes, kL2EvictFirst);
        else {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + blk * kWarpObsBytes),
                         "r"(smem_addr(stage[warp])), "r"((uint32_t)kWarpObsBytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        bulk_store_wait_read();
    }
}

// ---- memoised trajectory tables ------------------------------------------------------------------
// One thread per table entry runs the very simulation the step kernels would run (same code, same
// warp-collective form; trailing lanes of the last warp recompute the last entry and do not store).
__global__ void __launch_bounds__(256) pz_build_land_table_kernel(uint16_t *tab) {
    const int64_t total = kTabLandEntries;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool store = idx < total;
    if (!store) idx = total - 1;
    const int x = (int)(idx % kTabNx);
    int64_t r = idx / kTabNx;
    const int y = (int)(r % kTabNy);
    r /= kTabNy;
    const int xv = (int)(r % kTabNxv) - 20;
    const int yv = (int)(r / kTabNxv) - kTabYv;
    bool g;
    const int lx = simulate_landing_x<false>(kFullMask, x, y, xv, yv, true, g);
    if (store) tab[idx] = (uint16_t)((unsigned)lx | (g ? 0x8000u : 0u));
}

__global__ void __launch_bounds__(256) pz_build_power_table_kernel(uint16_t *tab) {
    const int64_t total = kTabPowerEntries;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool store = idx < total;
    if (!store) idx = total - 1;
    const int x = (int)(idx % kTabNx);
    int64_t r = idx / kTabNx;
    const int y = (int)(r % kTabNy);
    r /= kTabNy;
    const int half_yv0 = (int)(r % kTabNyv) - kTabYv;
    const int xd = (int)(r / kTabNyv);
    const int xv0 = (x < kGroundHalfWidth) ? (xd + 1) * 10 : -(xd + 1) * 10;  // physics.py:841-844
    bool g;
    const int lx = simulate_landing_x<true>(kFullMask, x, y, xv0, 2 * half_yv0, true, g);
    if (store) tab[idx] = (uint16_t)lx;  // the landing x alone: the search (computer_decide) has no use for `g`
}

struct DeviceTables {
    uint16_t *land = nullptr, *power = nullptr;
    int state = 0;    // 0 = not built, 1 = ready, -1 = failed (the kernels iterate instead)
    int error = 0;    // why: the cudaError_t of the failed allocation / build
    std::mutex mu;    // per device: ranks-as-threads on different GPUs build concurrently
};
constexpr int kMaxDevices = 64;
static DeviceTables g_tables[kMaxDevices];

// Returns the current device's tables, building them on first use (synchronises `stream` once; under stream
// capture nothing can be built: call pz_tables_prepare beforehand). *err receives the reason when there are none.
static const DeviceTables *acquire_tables(cudaStream_t stream, int *err = nullptr) {
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        if (err) *err = e0 != cudaSuccess ? (int)e0 : (int)cudaErrorInvalidDevice;
        return nullptr;
    }
    DeviceTables &t = g_tables[dev];
    std::lock_guard<std::mutex> lock(t.mu);
    if (t.state == 0) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
            if (err) *err = (int)cudaErrorStreamCaptureUnsupported;
            return nullptr;  // not built and cannot be built now: this launch iterates; state stays 0
        }
        cudaError_t e1 = cudaMalloc(&t.land, kTabLandEntries * sizeof(uint16_t));
        cudaError_t e2 = e1 == cudaSuccess ? cudaMalloc(&t.power, kTabPowerEntries * sizeof(uint16_t)) : e1;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            if (t.land) cudaFree(t.land);
            t.land = t.power = nullptr;
            t.state = -1;
            t.error = (int)(e1 != cudaSuccess ? e1 : e2);
            cudaGetLastError();  // clear the sticky allocation error: the iterative path needs no table
        } else {
            pz_build_land_table_kernel<<<(unsigned)((kTabLandEntries + 255) / 256), 256, 0, stream>>>(t.land);
            pz_build_power_table_kernel<<<(unsigned)((kTabPowerEntries + 255) / 256), 256, 0, stream>>>(t.power);
            cudaError_t e3 = cudaStreamSynchronize(stream);
            if (e3 == cudaSuccess) e3 = cudaGetLastError();
            if (e3 != cudaSuccess) {
                cudaFree(t.land);
                cudaFree(t.power);
                t.land = t.power = nullptr;
                t.state = -1;
                t.error = (int)e3;
            } else {
                t.state = 1;
            }
        }
    }
    if (t.state != 1 && err) *err = t.error ? t.error : (int)cudaErrorUnknown;
    return t.state == 1 ? &t : nullptr;
}

// ---- host side ---------------------------------------------------------------------------------
int check_config(const pz_config *c) {
    if (!c) return PZ_E_BADARG;
    // ABI handshake first: nothing past the two leading words is read from a struct of another revision
    if (c->struct_bytes != (uint32_t)sizeof(pz_config) || c->abi_version != (uint32_t)PZ_VERSION) return PZ_E_ABI;
    if (c->winning_score < 1 || c->winning_score > 1023) return PZ_E_BADCONFIG;
    if (c->serve < 0 || c->serve > 2) return PZ_E_BADCONFIG;
    if (c->action_dtype < 0 || c->action_dtype > 2) return PZ_E_BADCONFIG;
    if (c->reward_dtype < 0 || c->reward_dtype > 1) return PZ_E_BADCONFIG;
    if (c->obs_dtype < PZ_OBS_I32 || c->obs_dtype > PZ_OBS_F64) return PZ_E_BADCONFIG;
    if (c->normalize_observation && (c->obs_dtype == PZ_OBS_I32 || c->obs_dtype == PZ_OBS_I16)) return PZ_E_BADCONFIG;
    if (c->reward_in_normal_state < PZ_RINS_OFF || c->reward_in_normal_state > PZ_RINS_INNER) return PZ_E_BADCONFIG;
    if (c->max_episode_frames < 0) return PZ_E_BADCONFIG;
    if (c->obs_layout < PZ_LAYOUT_ENV_MAJOR || c->obs_layout > PZ_LAYOUT_ENV_MAJOR_SHARED) return PZ_E_BADCONFIG;
    if (c->obs_layout == PZ_LAYOUT_ENV_MAJOR_SHARED && c->obs_dtype != PZ_OBS_I32 && c->obs_dtype != PZ_OBS_I16)
        return PZ_E_BADCONFIG;  // shared rows are the integer observation
    if (c->obs_layout == PZ_LAYOUT_FEATURE_MAJOR && c->obs_feature_rows != 0 && c->obs_feature_rows < PZ_OBS_WORDS)
        return PZ_E_BADCONFIG;
    return 0;
}

static void fill_params(KParams &P, int32_t *state, int64_t n, const pz_config *c) {
    memset(&P, 0, sizeof(P));
    P.state = state;
    P.n = n;
    P.begin = 0;
    P.end = n;
    P.cfg.winning_score = c->winning_score;
    P.cfg.serve = c->serve;
    P.autoreset = c->autoreset != 0;
    P.simplify = c->simplify_action != 0;
    P.shaped = c->reward_by_ball_position != 0 || c->reward_in_normal_state != PZ_RINS_OFF;
    P.act_dtype = c->action_dtype;
    P.rew_dtype = c->reward_dtype;
    P.obs_dtype = c->obs_dtype;
    P.normalize = c->normalize_observation != 0;
    P.max_frames = c->max_episode_frames;
    P.obs_layout = c->obs_layout;
    P.obs_rows = c->obs_feature_rows > 0 ? c->obs_feature_rows : PZ_OBS_WORDS;
    // outputs are written once and never read back by the simulator: evict them from L2 first
    // (-1.7 us per 1 M-env launch). The state words keep the normal policy: evict-last on them, meant to
    // hold them in the 126 MB L2 across launches, measured 2 us SLOWER.
    P.pdl = (c->flags & PZ_FLAG_NO_PDL) ? 0 : 1;
    P.state_policy = kL2EvictNormal;
    P.out_policy = (c->flags & PZ_FLAG_NO_L2_HINTS) ? kL2EvictNormal : kL2EvictFirst;
    P.x_line = c->x_line;
    P.y_line = c->y_line;
    for (int agent = 0; agent < 2; agent++)
        for (int b = 0; b < 3; b++)
            for (int z = 0; z < 4; z++) {
                // the wrapper stack evaluated in double, as Python does on `int` / `float` rewards, innermost
                // wrapper first: reward_in_normal_state.py:13-14 `if rews[agent] == 0: rews[agent] = reward`,
                // reward_by_ball_position.py:28-29 `rews[agent] += additional_reward[agent*4 + zone]`
                double r = (double)(b - 1);
                if (c->reward_in_normal_state == PZ_RINS_INNER && r == 0.0) r = c->normal_state_reward;
                if (c->reward_by_ball_position) r = r + c->additional_reward[agent * 4 + z];
                if (c->reward_in_normal_state == PZ_RINS_OUTER && r == 0.0) r = c->normal_state_reward;
                P.table[agent * 12 + b * 4 + z] = r;
            }
}

// Computer players read the memoised trajectory tables unless PZ_FLAG_NO_TABLES is set.
static void attach_tables(KParams &P, const pz_config *c, cudaStream_t stream) {
    if ((c->is_player1_computer || c->is_player2_computer) && !(c->flags & PZ_FLAG_NO_TABLES)) {
        if (const DeviceTables *t = acquire_tables(stream)) {
            P.cfg.tab_land = t->land;
            P.cfg.tab_power = t->power;
        }
    }
}

static inline int ai_mask(const pz_config *c) {
    return (c->is_player1_computer ? 1 : 0) | (c->is_player2_computer ? 2 : 0);
}

static inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int launch_status() {
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : (int)err;
}

int launch_reset(int32_t *state_dev, int64_t n, int64_t begin, int64_t end, const pz_config *cfg, void *obs_dev,
                 const pz_episode_io *ep, cudaStream_t stream) {
    if (!state_dev || n < 0 || begin < 0 || end > n || begin > end || (begin & 31)) return PZ_E_BADARG;
    if (int rc = check_config(cfg)) return rc;
    if (!aligned16(state_dev) || !aligned16(obs_dev)) return PZ_E_ALIGN;
    if (end == begin) return 0;
    KParams P;
    fill_params(P, state_dev, n, cfg);
    P.begin = begin;
    P.end = end;
    P.obs = obs_dev;
    if (ep) {
        if (!aligned16(ep->episode_return_dev)) return PZ_E_ALIGN;
        P.ep_return = reinterpret_cast<double2 *>(ep->episode_return_dev);
        P.ep_length = ep->episode_length_dev;
    }
    pz_reset_kernel<<<grid_for(end - begin), kThreads, 0, stream>>>(P);
    return launch_status();
}

int launch_observe(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, cudaStream_t stream) {
    if (!state_dev || !obs_dev || n < 0) return PZ_E_BADARG;
    if (int rc = check_config(cfg)) return rc;
    if (!aligned16(state_dev) || !aligned16(obs_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    KParams P;
    fill_params(P, state_dev, n, cfg);
    P.obs = obs_dev;
    pz_observe_kernel<<<grid_for(n), kThreads, 0, stream>>>(P);
    return launch_status();
}

int launch_step(int32_t *state_dev, int64_t n, int64_t begin, int64_t end, const pz_config *cfg,
                const void *actions_dev, void *obs_dev, void *reward_dev, uint8_t *done_dev, int64_t *stats_dev,
                const pz_episode_io *ep, cudaStream_t st) {
    if (!state_dev || n < 0 || begin < 0 || end > n || begin > end || (begin & 31)) return PZ_E_BADARG;
    if (int rc = check_config(cfg)) return rc;
    const int am = ai_mask(cfg);
    if (!actions_dev && am != 3) return PZ_E_BADARG;  // actions may be omitted only when both play themselves
    if (!aligned16(state_dev) || !aligned16(obs_dev) || !aligned16(actions_dev) || !aligned16(reward_dev))
        return PZ_E_ALIGN;
    if (end == begin) return 0;
    KParams P;
    fill_params(P, state_dev, n, cfg);
    attach_tables(P, cfg, st);
    P.begin = begin;
    P.end = end;
    P.actions = actions_dev;
    P.obs = obs_dev;
    P.reward = reward_dev;
    P.done = done_dev;
    P.stats = reinterpret_cast<unsigned long long *>(stats_dev);
    if (ep) {
        if (!aligned16(ep->episode_return_dev)) return PZ_E_ALIGN;
        P.ep_return = reinterpret_cast<double2 *>(ep->episode_return_dev);
        P.ep_length = ep->episode_length_dev;
        P.truncated = ep->truncated_dev;
        P.status = ep->status_dev;
        if (ep->seq_dev) {
            if (end - begin > kThreads) return PZ_E_BADARG;  // the completion word needs a single CTA
            P.seq = ep->seq_dev;
            P.seq_value = ep->seq_value;
        }
    }
    switch (am) {
        case 0: launch_step_kernel<0>(end - begin, st, P); break;
        case 1: launch_step_kernel<1>(end - begin, st, P); break;
        case 2: launch_step_kernel<2>(end - begin, st, P); break;
        default: launch_step_kernel<3>(end - begin, st, P); break;
    }
    return launch_status();
}

}  // namespace pz

using namespace pz;

extern "C" {

int pz_version(void) { return PZ_VERSION; }
int pz_state_words(void) { return PZ_STATE_WORDS; }
int pz_unpacked_words(void) { return PZ_UNPACKED_WORDS; }
size_t pz_state_bytes(int64_t n) { return n < 0 ? 0 : (size_t)n * PZ_STATE_WORDS * sizeof(int32_t); }

const char *pz_strerror(int code) {
    switch (code) {
        case 0: return "success";
        case PZ_E_BADARG: return "pikazoo_b200: bad argument";
        case PZ_E_BADCONFIG: return "pikazoo_b200: bad config (winning_score must be in [1,1023]; serve/dtype codes)";
        case PZ_E_ALIGN: return "pikazoo_b200: state/obs pointers must be 16-byte aligned";
        case PZ_E_NODEVICE: return "pikazoo_b200: no usable sm_100 device";
        case PZ_E_ABI:
            return "pikazoo_b200: pz_config.struct_bytes / abi_version do not match this library (the caller's binding "
                   "declares another revision of struct pz_config; initialise it with pz_config_init)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "pikazoo_b200: unknown error";
    }
}

size_t pz_config_bytes(void) { return sizeof(pz_config); }

int pz_obs_player2_index(int k) {
    if (k < 0 || k >= PZ_OBS_WORDS) return -1;
    return k < 13 ? k + 13 : (k < 26 ? k - 13 : k);
}

int pz_config_init(pz_config *c, size_t caller_struct_bytes) {
    if (!c) return PZ_E_BADARG;
    if (caller_struct_bytes != sizeof(pz_config)) return PZ_E_ABI;  // nothing is written to a struct of another size
    memset(c, 0, sizeof(*c));
    c->struct_bytes = (uint32_t)sizeof(pz_config);
    c->abi_version = (uint32_t)PZ_VERSION;
    c->winning_score = 15;
    c->serve = PZ_SERVE_WINNER;
    c->x_line = 216;
    c->y_line = 176;
    c->autoreset = 1;
    return 0;
}

int pz_seed(int32_t *state_dev, int64_t n, uint64_t base_seed, uint64_t first_env, void *stream) {
    if (!state_dev || n < 0) return PZ_E_BADARG;
    if (!aligned16(state_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    pz_seed_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state_dev, n, base_seed, first_env, nullptr);
    return launch_status();
}

int pz_seed_array(int32_t *state_dev, int64_t n, const uint64_t *seeds_dev, void *stream) {
    if (!state_dev || !seeds_dev || n < 0) return PZ_E_BADARG;
    if (!aligned16(state_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    pz_seed_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state_dev, n, 0, 0, seeds_dev);
    return launch_status();
}

size_t pz_obs_elem_bytes(int32_t obs_dtype) {
    switch (obs_dtype) {
        case PZ_OBS_I32:
        case PZ_OBS_F32: return 4;
        case PZ_OBS_I16:
        case PZ_OBS_F16:
        case PZ_OBS_BF16: return 2;
        case PZ_OBS_F64: return 8;
        default: return 0;
    }
}

int pz_reset(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, void *stream) {
    return pz::launch_reset(state_dev, n, 0, n, cfg, obs_dev, nullptr, (cudaStream_t)stream);
}

int pz_observe(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, void *stream) {
    return pz::launch_observe(state_dev, n, cfg, obs_dev, (cudaStream_t)stream);
}

int pz_reset_ex(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, const pz_episode_io *episode,
                void *stream) {
    return pz::launch_reset(state_dev, n, 0, n, cfg, obs_dev, episode, (cudaStream_t)stream);
}

int pz_step(int32_t *state_dev, int64_t n, const pz_config *cfg, const void *actions_dev, void *obs_dev,
            void *reward_dev, uint8_t *done_dev, int64_t *stats_dev, void *stream) {
    return pz::launch_step(state_dev, n, 0, n, cfg, actions_dev, obs_dev, reward_dev, done_dev, stats_dev, nullptr,
                           (cudaStream_t)stream);
}

int pz_step_ex(int32_t *state_dev, int64_t n, const pz_config *cfg, const void *actions_dev, void *obs_dev,
               void *reward_dev, uint8_t *done_dev, int64_t *stats_dev, const pz_episode_io *episode, void *stream) {
    return pz::launch_step(state_dev, n, 0, n, cfg, actions_dev, obs_dev, reward_dev, done_dev, stats_dev, episode,
                           (cudaStream_t)stream);
}

int pz_rollout(int32_t *state_dev, int64_t n, const pz_config *cfg, int32_t K, int32_t action_source,
               uint64_t action_seed, uint64_t first_env, uint64_t frame0, void *obs_dev, int64_t *stats_dev,
               void *stream) {
    if (!state_dev || n < 0 || K < 1) return PZ_E_BADARG;
    if (action_source != PZ_ACTIONS_NOOP && action_source != PZ_ACTIONS_SYNTH) return PZ_E_BADARG;
    if (int rc = check_config(cfg)) return rc;
    if (!aligned16(state_dev) || !aligned16(obs_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    KParams P;
    fill_params(P, state_dev, n, cfg);
    attach_tables(P, cfg, (cudaStream_t)stream);
    P.obs = obs_dev;
    P.stats = reinterpret_cast<unsigned long long *>(stats_dev);
    P.K = K;
    P.action_source = action_source;
    P.action_seed = action_seed;
    P.first_env = first_env;
    P.frame0 = frame0;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned rollout_grid = (unsigned)((n + kRolloutThreads - 1) / kRolloutThreads);
    const bool plain = action_source == PZ_ACTIONS_NOOP && P.max_frames <= 0;
#define PZ_ROLLOUT_CASE(M)                                                                      \
    case M:                                                                                     \
        if (plain && M != 0 && P.cfg.tab_land != nullptr && P.cfg.tab_power != nullptr)         \
            pz_rollout_kernel<M, true, M != 0><<<rollout_grid, kRolloutThreads, 0, st>>>(P);    \
        else if (plain)                                                                         \
            pz_rollout_kernel<M, true><<<rollout_grid, kRolloutThreads, 0, st>>>(P);            \
        else                                                                                    \
            pz_rollout_kernel<M, false><<<rollout_grid, kRolloutThreads, 0, st>>>(P);           \
        break;
    switch (ai_mask(cfg)) {
        PZ_ROLLOUT_CASE(0)
        PZ_ROLLOUT_CASE(1)
        PZ_ROLLOUT_CASE(2)
        PZ_ROLLOUT_CASE(3)
    }
#undef PZ_ROLLOUT_CASE
    return launch_status();
}

int pz_tables_prepare(void *stream) {
    int err = 0;
    return acquire_tables((cudaStream_t)stream, &err) ? 0 : err;
}

size_t pz_tables_bytes(void) { return (size_t)(kTabLandEntries + kTabPowerEntries) * sizeof(uint16_t); }

int pz_tables_ready(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    std::lock_guard<std::mutex> lock(g_tables[dev].mu);
    return g_tables[dev].state == 1;
}

void pz_tables_release(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    DeviceTables &t = g_tables[dev];
    std::lock_guard<std::mutex> lock(t.mu);
    if (t.state == 1) {
        cudaDeviceSynchronize();
        cudaFree(t.land);
        cudaFree(t.power);
    }
    t.land = t.power = nullptr;
    t.state = 0;
    t.error = 0;
}

int pz_probe_write(void *dst_dev, size_t bytes, int32_t mode, void *stream) {
    if (!dst_dev || (bytes & 15u) || mode < 0 || mode > 4) return PZ_E_BADARG;
    if (!aligned16(dst_dev)) return PZ_E_ALIGN;
    if (bytes == 0) return 0;
    if (mode <= 2) {
        const size_t n16 = bytes / 16;
        pz_probe_write_kernel<<<(unsigned)((n16 + 1023) / 1024), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<uint4 *>(dst_dev), n16, mode, 0x5A5A5A5Au);
    } else {
        const size_t blocks = bytes / kWarpObsBytes;  // whole 8,960-byte blocks only
        if (blocks == 0) return PZ_E_BADARG;
        pz_probe_bulk_kernel<<<(unsigned)((blocks + kWarps - 1) / kWarps), kThreads, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<char *>(dst_dev), blocks, mode == 3, 0x5A5A5A5Au);
    }
    return launch_status();
}

int pz_export_state(const int32_t *state_dev, int64_t n, int32_t *unpacked_dev, void *stream) {
    if (!state_dev || !unpacked_dev || n < 0) return PZ_E_BADARG;
    if (!aligned16(state_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    pz_export_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state_dev, n, unpacked_dev);
    return launch_status();
}

int pz_import_state(int32_t *state_dev, int64_t n, const int32_t *unpacked_dev, void *stream) {
    if (!state_dev || !unpacked_dev || n < 0) return PZ_E_BADARG;
    if (!aligned16(state_dev)) return PZ_E_ALIGN;
    if (n == 0) return 0;
    pz_import_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(state_dev, n, unpacked_dev);
    return launch_status();
}

}  // extern "C"
