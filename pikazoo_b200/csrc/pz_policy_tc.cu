// pz_policy_mlp_act on the 5th-generation tensor cores (tcgen05 + TMEM), tiles by TMA.
//
// Thread t of a 128-env tile IS env t from the first epilogue on, which is what makes this form cheap: in the
// warp-level mma.sync kernel (pz_policy.cu) a warp's 16 envs are spread over accumulator fragments, so every
// hidden activation is rectified / packed and the sample is taken in fragment order (24 candidate slots per quad
// for 18 actions, quad shuffles), ~1,180 warp instructions per 16 envs. Here, per (tile, agent):
//
//   tile      five cp.async.bulk.tensor.3d (one per group of 8 features) bring the observation tile from the
//             feature-major tensor into the canonical layout below; completion on an mbarrier
//   layer 1   D1[128 envs][80 hidden]  = X^T[128][48] . W1^T      3 tcgen05.mma (K = 16 each), one issuing thread;
//             (TMEM, fp32, 80 columns)                            A = the tile (MN-major, no swizzle), B = W1
//                                                                 (K-major), both in shared memory
//   epilogue  thread t reads row t of D1 (tcgen05.ld 32x32b), rectifies, rounds to bf16 and stores the 40 packed
//             pairs back over the columns it has read (tcgen05.st) as row t of H
//   layer 2   D2[128][32]              = H[128][80] . W2^T        5 tcgen05.mma with A taken from TMEM
//   sample    thread t reads its env's logits (one row of D2) and inverts the cumulative distribution in registers
//             (sample_inverse_cdf, pz_policy.cuh): no shuffles, no padded slots
//
// The tensor-core work of one (tile, agent) is a serial chain (MMA -> epilogue -> MMA -> sample) of latencies, and
// a chain holds 80 TMEM columns while it runs (D1 0..79, H in place 0..39, D2 40..71), so the SM's 512 columns
// carry six of them: ONE persistent CTA per SM stages the weights once and runs six chains of 128 threads, two per
// tile (one per agent), each with its own named barrier, mbarriers, TMEM columns and a double-buffered observation
// tile: the tile after next is fetched while a tile is computed (its bulk copies are issued while layer 2 is in the
// tensor core), and layer 1 of the NEXT tile is issued as soon as every thread holds its logits in registers, so
// that its round trip through the tensor core hides behind the sampling arithmetic. Measured on B200, 2 M envs:
// 0.094 ms (the mma.sync kernel 0.219; the steps in between are in DESIGN.md section 4a).
// PZ_TC_AGENTS_PER_CHAIN=2 builds three chains of 256 threads that batch both agents' MMAs; PZ_TC_TIMING a debug
// build that stamps a chain's phases (profiles/policy_chain_phases.py).
//
// Canonical shared-memory layouts (no swizzle; 8 x 16-byte "core matrices" of 128 contiguous bytes):
//   X  (MN-major A): element (env m, feature k) at (k / 8) * 2048 + (m / 8) * 128 + (k % 8) * 16 + (m % 8) * 2
//                    -> a 16-byte piece of 8 consecutive envs of one feature row lands as one 16-byte row of a
//                       core matrix: an (8 envs, 8 features, 16 env atoms) TMA box is one group of 8 features
//   W1 (K-major B):  element (hidden n, feature k) at (k / 8) * 1280 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2
//   W2 (K-major B):  element (action n, hidden k)  at (k / 8) *  512 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2
#include <cuda.h>  // CUtensorMap; the driver entry point is looked up at run time, libcuda is not linked
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <mutex>

#include "pz_policy.cuh"
#include "pz_tcgen05.cuh"

namespace pzp {
namespace tc {

#ifndef PZ_TC_AGENTS_PER_CHAIN
#define PZ_TC_AGENTS_PER_CHAIN 1
#endif
constexpr int kAPC = PZ_TC_AGENTS_PER_CHAIN;   // 1: a chain of 4 warps owns one agent of a tile;
                                               // 2: a chain of 8 warps owns a tile (warps 0-3 agent 0, 4-7 agent 1)
constexpr int kChains = 6 / kAPC;              // per CTA = per SM
constexpr int kChainThreads = 128 * kAPC;      // env row = 32 * (warp % 4) + lane
constexpr int kThreads = kChains * kChainThreads;
constexpr int kTileEnvs = 128;
constexpr int kKP = PZ_POLICY_MAX_FEATURES;  // 48 = 3 k-steps of 16
constexpr int kHP = PZ_POLICY_MAX_HIDDEN;    // 80: N of layer 1 (multiple of 16), 5 k-steps of layer 2
constexpr int kAP = 32;                      // N of layer 2: PZ_POLICY_MAX_ACTIONS (24) padded to a multiple of 16
constexpr uint32_t kTmemCols = 512;          // the whole SM
constexpr uint32_t kAgentCols = kHP, kChainCols = kAPC * kAgentCols;
constexpr uint32_t kColD1 = 0, kColH = 0, kColD2 = kHP / 2;  // H overwrites the columns of D1 its thread has read
static_assert(kChains * kChainCols <= kTmemCols && kHP / 2 + kAP <= kHP, "TMEM columns");
static_assert(PZ_POLICY_MAX_ACTIONS <= kAP && kKP % 16 == 0 && kHP % 16 == 0, "tile shapes");

constexpr int kXKGroup = (kTileEnvs / 8) * 128;    // 2048 B: one group of 8 features, 16 env atoms
constexpr int kXAgent = (kKP / 8) * kXKGroup;      // 12288 B
constexpr int kXTile = kAPC * kXAgent;             // what one chain stages: both agents, or its own
constexpr int kW1KGroup = (kHP / 8) * 128;         // 1280 B
constexpr int kW1Agent = (kKP / 8) * kW1KGroup;    // 7680 B
constexpr int kW2KGroup = (kAP / 8) * 128;         // 512 B
constexpr int kW2Agent = (kHP / 8) * kW2KGroup;    // 5120 B
constexpr int kOffW1 = 0, kOffW2 = 2 * kW1Agent, kOffX = kOffW2 + 2 * kW2Agent;  // X: [chain][buffer][agent]
constexpr int kOffBar = kOffX + kChains * 2 * kXTile;
constexpr int kOffTmemSlot = kOffBar + 24 * kChains;
constexpr size_t kSmemBytes = kOffTmemSlot + 16;   // per chain: MMA mbarrier + one "tile landed" mbarrier per buffer
                                                    // (8 B each); the TMEM base address (4 B)

#ifndef PZ_TC_EPI_BATCH
#define PZ_TC_EPI_BATCH 2
#endif
__device__ __forceinline__ void chain_sync(int chain) {
    asm volatile("bar.sync %0, %1;" ::"r"(chain + 1), "r"(kChainThreads) : "memory");
}

#ifdef PZ_TC_TIMING  // debug build: chain 0 of CTA 0 stamps its phases into the logits buffer (as uint64)
#define PZ_STAMP(slot)                                                                                     \
    do {                                                                                                   \
        if (blockIdx.x == 0 && tid == 0 && iter >= 4 && iter < 12)                                         \
            reinterpret_cast<long long *>(P.logits)[(iter - 4) * 12 + (slot)] = clock64();                 \
    } while (0)
#else
#define PZ_STAMP(slot) do { } while (0)
#endif

template <int NA>
__global__ void __launch_bounds__(kThreads, 1)
    pz_policy_mlp_tc_kernel(const __grid_constant__ Params P, const __grid_constant__ CUtensorMap tmap, const int use_tma) {
    const int n_actions = NA ? NA : P.n_actions;
    constexpr int kCand = NA ? NA : PZ_POLICY_MAX_ACTIONS;  // candidates the unrolled loops run over
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kOffTmemSlot);
    const int tid = threadIdx.x, lane = tid & 31;
    // warp-uniform by construction: everything derived from it (chain, agent, TMEM and shared-memory addresses, the
    // MMA descriptors) can live in uniform registers, which is where the tensor-core instructions take them from
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
    const int chain = warp / (4 * kAPC), wic = warp % (4 * kAPC), ctid = tid - chain * kChainThreads;
    const int agent = kAPC == 2 ? wic >> 2 : chain & 1;   // the agent this thread samples for
    const int slot = kAPC == 2 ? agent : 0;               // its place in the chain's TMEM columns and tile buffers
    const int seq = kAPC == 2 ? chain : chain >> 1;       // which of the CTA's three tile sequences the chain walks
    const uint32_t bar = s_base + kOffBar + 24 * chain;  // +8, +16: the tile buffers' barriers

    // Programmatic dependent launch (both calls are no-ops for an ordinary launch): the next kernel in the stream
    // — the step kernel, which parks in its own cudaGridDependencySynchronize — may be scheduled as this grid's CTAs
    // leave; and this kernel's set-up below (TMEM, barriers, 173 KB of zeroed shared memory) runs while the previous
    // kernel's last CTAs drain. Nothing that a predecessor may have written (weights, observations) is read before
    // cudaGridDependencySynchronize.
    cudaTriggerProgrammaticLaunchCompletion();
    // ---- once per CTA: TMEM, the barriers, zeroed operands, the weights in canonical K-major order ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kOffTmemSlot),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (ctid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + 8), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar + 16), "r"(1u) : "memory");
    }
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = tid; i < kOffBar / 16; i += kThreads) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    cudaGridDependencySynchronize();
    {
        __nv_bfloat16 *w1s = reinterpret_cast<__nv_bfloat16 *>(smem + kOffW1);
        for (int i = tid; i < 2 * P.h1 * P.k1; i += kThreads) {
            const int a = i / (P.h1 * P.k1), rem = i - a * (P.h1 * P.k1), n = rem / P.k1, k = rem - n * P.k1;
            w1s[(a * kW1Agent + (k >> 3) * kW1KGroup + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = P.w1[i];
        }
        __nv_bfloat16 *w2s = reinterpret_cast<__nv_bfloat16 *>(smem + kOffW2);
        for (int i = tid; i < 2 * n_actions * P.k2; i += kThreads) {
            const int a = i / (n_actions * P.k2), rem = i - a * (n_actions * P.k2), n = rem / P.k2, k = rem - n * P.k2;
            w2s[(a * kW2Agent + (k >> 3) * kW2KGroup + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) >> 1] = P.w2[i];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_chain = tmem + chain * kChainCols;                                       // lane 0: MMA operands
    const uint32_t t_row = t_chain + slot * kAgentCols + ((uint32_t)((wic & 3) * 32) << 16);  // this thread's row
    const uint32_t s_x = s_base + kOffX + chain * 2 * kXTile;

    constexpr uint32_t kIdesc1 = instr_desc(kHP, true), kIdesc2 = instr_desc(kAP, false);
    const bool vec_ok = (P.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(P.obs) & 15u) == 0);
    const int64_t n_tiles = (P.n + kTileEnvs - 1) / kTileEnvs;
    const int64_t t_stride = (int64_t)gridDim.x * 3;

    // The chain's threads fill what it computes on (its agent's half of the tile, or both): per pass one group
    // of 8 features, lane = (feature % 8) + 8 * (env atom % 4), warp % 4 = env atom / 4 — a warp writes 512
    // contiguous bytes of shared memory and reads 8 rows x 64 contiguous bytes.
    auto load_tile = [&](int64_t t, uint32_t b) {
        const int64_t env0 = t * kTileEnvs;
        if (use_tma) {
            // One bulk tensor copy per group of 8 features: the tensor map views the observations as
            // [env / 8][feature row][env % 8] with an (8, 8, 16) box, whose dense shared-memory image (8 envs, then
            // 8 features, then 16 env atoms) IS the canonical layout of the group. Envs past the end read as zero.
            if (wic == 0 && elect_one()) {
                const uint32_t full = bar + 8 + 8 * b;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                             "r"((uint32_t)(kAPC * ((P.k1 + 7) >> 3) * kXKGroup))
                             : "memory");
                for (int sl = 0; sl < kAPC; sl++)
                    for (int kg = 0; kg < ((P.k1 + 7) >> 3); kg++)
                        asm volatile(
                            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                            ::"r"(s_x + b * kXTile + sl * kXAgent + kg * kXKGroup), "l"(&tmap), "r"(0),
                            "r"((kAPC == 2 ? sl : agent) * P.rows + kg * 8), "r"((int)(env0 >> 3)), "r"(full)
                            : "memory");
            }
            return;
        }
        if (vec_ok && env0 + kTileEnvs <= P.n) {
            const int kr = lane & 7, m8 = (wic & 3) * 4 + (lane >> 3);
            const __nv_bfloat16 *src = P.obs + ((int64_t)agent * P.rows + kr) * P.ld + env0 + m8 * 8;
            uint32_t dst = s_x + b * kXTile + slot * kXAgent + m8 * 128 + kr * 16;
            for (int k = kr; k < P.k1; k += 8, src += 8 * P.ld, dst += kXKGroup)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        } else {  // ragged last tile or unaligned rows
            __nv_bfloat16 *xs = reinterpret_cast<__nv_bfloat16 *>(smem + kOffX + (chain * 2 + b) * kXTile);
            const __nv_bfloat16 zero = __float2bfloat16(0.0f);
            for (int i = ctid; i < kAPC * P.k1 * kTileEnvs; i += kChainThreads) {
                const int sl = i / (P.k1 * kTileEnvs), rem = i - sl * (P.k1 * kTileEnvs), k = rem / kTileEnvs,
                          m = rem % kTileEnvs, a = kAPC == 2 ? sl : agent;
                xs[(sl * kXAgent + (k >> 3) * kXKGroup + (m >> 3) * 128 + (k & 7) * 16 + (m & 7) * 2) >> 1] =
                    env0 + m < P.n ? P.obs[((int64_t)a * P.rows + k) * P.ld + env0 + m] : zero;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // layer 1 of the chain's agent(s) of the tile in buffer b -> D1 (one thread)
    uint32_t full_phase = 0;  // bit b: parity the next wait on buffer b's barrier uses (tracked by every thread)
    auto issue_layer1 = [&](uint32_t b) {
        if (use_tma) mbar_wait(bar + 8 + 8 * b, (full_phase >> b) & 1u);  // the bulk copies of this tile have landed
        tc_fence_after();
#pragma unroll
        for (int sl = 0; sl < kAPC; sl++)
#pragma unroll
            for (int ks = 0; ks < kKP / 16; ks++)
                mma_ss(t_chain + sl * kAgentCols + kColD1,
                       smem_desc(s_x + b * kXTile + sl * kXAgent + ks * 2 * kXKGroup, kXKGroup, 128),
                       smem_desc(s_base + kOffW1 + (kAPC == 2 ? sl : agent) * kW1Agent + ks * 2 * kW1KGroup, kW1KGroup, 128),
                       kIdesc1, ks > 0);
        mma_commit(bar);
    };
    // The chain's barrier before layer 1 of a tile is issued: every thread's TMEM reads of the previous tile are
    // done; without TMA, also: the cp.async data of every group but the newest has landed and is made visible to the
    // tensor core (with TMA the issuing thread waits on the tile's own mbarrier instead)
    auto tile_ready = [&]() {
        if (!use_tma) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        tc_fence_before();  // and this thread's TMEM reads are ordered before the MMAs issued after the barrier
        chain_sync(chain);
    };

    // Software pipeline over the chain's tiles t0, t1 = t0 + stride, ...: the tile after next is fetched while a
    // tile is computed, and layer 1 of the NEXT tile is issued as soon as every thread holds this tile's logits in
    // registers, so its round trip through the tensor core hides behind the sampling arithmetic.
    int64_t tile = blockIdx.x + (int64_t)gridDim.x * seq;
    uint32_t buf = 0, phase = 0;
    if (tile < n_tiles) {
        load_tile(tile, 0);
        if (tile + t_stride < n_tiles)
            load_tile(tile + t_stride, 1);
        else
            asm volatile("cp.async.commit_group;" ::: "memory");
        tile_ready();
        if (wic == 0 && elect_one()) issue_layer1(0);
        full_phase ^= 1u;
    }
    int iter = 0;
    for (; tile < n_tiles; tile += t_stride, buf ^= 1u, iter++) {
        PZ_STAMP(0);
        const int64_t env = tile * kTileEnvs + (wic & 3) * 32 + lane;
        const uint32_t nbase = P.greedy ? 0u : noise_base(P.seed, P.step, P.first_env + (uint64_t)env);
        mbar_wait(bar, phase);  // layer 1 of this tile
        phase ^= 1u;
        tc_fence_after();
        PZ_STAMP(1);
        // ---- relu, round to bf16, back to TMEM as the A operand of layer 2 (thread = env row of its agent).
        //      H column j holds hidden units 2j, 2j+1: it overwrites D1 columns this thread has already read.
        {
            constexpr int kChunks = kHP / 16, kBatch = PZ_TC_EPI_BATCH;  // chunks of 16 columns read per wait
            uint32_t v[kBatch][16], h[8];
#pragma unroll
            for (int c = 0; c < kChunks; c += kBatch) {
#pragma unroll
                for (int q = 0; q < kBatch; q++)
                    if (c + q < kChunks) tmem_ld16(t_row + kColD1 + 16 * (c + q), v[q]);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < kBatch; q++)
                    if (c + q < kChunks) {
#pragma unroll
                        for (int j = 0; j < 8; j++) h[j] = relu_pack(v[q][2 * j], v[q][2 * j + 1]);
                        tmem_st8(t_row + kColH + 8 * (c + q), h);
                    }
            }
        }
        tmem_st_wait();
        PZ_STAMP(2);
        tc_fence_before();
        chain_sync(chain);
        PZ_STAMP(3);
        // ---- layer 2 -> D2
        if (wic == 0 && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int sl = 0; sl < kAPC; sl++)
#pragma unroll
                for (int j = 0; j < kHP / 16; j++)
                    mma_ts(t_chain + sl * kAgentCols + kColD2, t_chain + sl * kAgentCols + kColH + 8 * j,
                           smem_desc(s_base + kOffW2 + (kAPC == 2 ? sl : agent) * kW2Agent + j * 2 * kW2KGroup, kW2KGroup, 128),
                           kIdesc2, j > 0);
            mma_commit(bar);
        }
        PZ_STAMP(4);
        // layer 1 has read this tile's buffer: the tile after next can travel into it (issued here, off the
        // critical path between the two layers)
        if (tile + 2 * t_stride < n_tiles)
            load_tile(tile + 2 * t_stride, buf);
        else
            asm volatile("cp.async.commit_group;" ::: "memory");
        PZ_STAMP(5);
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        PZ_STAMP(6);
        // ---- this (env, agent)'s logits into registers; then the tensor core can have the columns back
        uint32_t lg[24];
        {
            uint32_t v16[16], v8[8];
            tmem_ld16(t_row + kColD2, v16);
            if (NA == 0 || NA > 16) tmem_ld8(t_row + kColD2 + 16, v8);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j++) lg[j] = v16[j];
#pragma unroll
            for (int j = 0; j < 8; j++) lg[16 + j] = (NA == 0 || NA > 16) ? v8[j] : 0u;
        }
        PZ_STAMP(7);
        if (tile + t_stride < n_tiles) {  // chain-uniform
            tile_ready();
            PZ_STAMP(8);
            if (wic == 0 && elect_one()) issue_layer1(buf ^ 1u);
            full_phase ^= 1u << (buf ^ 1u);
        }
        PZ_STAMP(9);
        // ---- the sample
#ifndef PZ_TC_TIMING
        if (P.logits != nullptr && env < P.n) {  // launch-uniform pointer
#pragma unroll
            for (int j = 0; j < kCand; j++)
                if (j < n_actions) P.logits[(env * 2 + agent) * n_actions + j] = __uint_as_float(lg[j]);
        }
#endif
        int act;
        if (P.greedy) {  // the packed-key arg-max both implementations share
            float best = pack_key(-INFINITY, 31);
#pragma unroll
            for (int j = 0; j < kCand; j++)
                if (j < n_actions) best = fmaxf(best, pack_key(__uint_as_float(lg[j]), j));
            act = 31 - (int)(__float_as_uint(best) & 31u);
            if (act >= n_actions) act = 0;  // every key NaN: action 0
        } else {
            float logit[kCand];
#pragma unroll
            for (int j = 0; j < kCand; j++) logit[j] = __uint_as_float(lg[j]);
            act = sample_inverse_cdf<kCand>(logit, n_actions, nbase + (uint32_t)(32 * agent) * 0x9E3779B9u);
        }
        if (env < P.n) {
            if (P.act_dtype == PZ_ACT_U8)
                reinterpret_cast<unsigned char *>(P.actions)[env * 2 + agent] = (unsigned char)act;
            else if (P.act_dtype == PZ_ACT_I32)
                reinterpret_cast<int *>(P.actions)[env * 2 + agent] = act;
            else
                reinterpret_cast<long long *>(P.actions)[env * 2 + agent] = act;
        }
        PZ_STAMP(10);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// The feature-major observations [2 * rows][ld] bf16 as a rank-3 tensor (env % 8, feature row, env / 8) with an
// (8, 8, 16) box: see load_tile. False when the shape does not allow it (the kernel then fills its tiles with
// cp.async / plain loads).
static bool make_obs_map(const Params &P, CUtensorMap *map) {
    if (P.ld % 8 != 0 || (reinterpret_cast<uintptr_t>(P.obs) & 15u) != 0 || P.k1 % 8 != 0 || P.rows < P.k1 || P.n < 8)
        return false;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {8, (cuuint64_t)(2 * (int64_t)P.rows), (cuuint64_t)(P.n / 8)};  // whole groups of 8 envs
    const cuuint64_t strides[2] = {(cuuint64_t)P.ld * 2, 16};                                    // bytes, dims 1 and 2
    const cuuint32_t box[3] = {8, 8, (cuuint32_t)(kTileEnvs / 8)}, estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16 *>(P.obs), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NA>
static cudaError_t launch_one(const Params &P, unsigned grid, cudaStream_t stream, int dev) {
    static std::mutex mu;
    static bool attr_set[64] = {};
    {  // more than 48 KB of dynamic shared memory needs the opt-in, once per device
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(pz_policy_mlp_tc_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kSmemBytes);
            if (e != cudaSuccess) return e;
            attr_set[dev] = true;
        }
    }
    alignas(64) CUtensorMap map;
    // whole groups of 8 envs only travel by bulk copy: a ragged tail (n % 8 != 0) keeps the element-wise path
    const int use_tma = (P.n % 8 == 0 && make_obs_map(P, &map)) ? 1 : 0;
    if (!use_tma) memset(&map, 0, sizeof(map));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, pz_policy_mlp_tc_kernel<NA>, P, map, use_tma);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace tc

int launch_tc(const Params &P, cudaStream_t stream) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // one persistent CTA per SM (it owns the SM's TMEM); chain c of CTA b walks tiles b + grid * (c + 3 i), so a
    // batch of fewer tiles than SMs still spreads one tile per SM
    const int64_t tiles = (P.n + tc::kTileEnvs - 1) / tc::kTileEnvs;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    cudaError_t err;
    if (P.n_actions == 18)
        err = tc::launch_one<18>(P, grid, stream, dev);
    else if (P.n_actions == 13)
        err = tc::launch_one<13>(P, grid, stream, dev);
    else
        err = tc::launch_one<0>(P, grid, stream, dev);
    return err == cudaSuccess ? 0 : (int)err;
}

}  // namespace pzp
