// Caller side of the hot path (BASELINE.json configs[4]: "actions from an on-device MLP policy in a rollout
// loop"): the two-layer MLP of pikazoo_b200/policy.py evaluated for both agents and sampled, in ONE kernel,
// straight from the simulator's feature-major bf16 observations.
//
//   logits[env][agent][:] = W2[agent] . relu(W1[agent] . x[agent][:, env])         bf16 in, fp32 accumulate
//   action = argmax_a (logits[a] + Gumbel noise(seed, step, global env, agent, a))   == a categorical sample
//            (on keys truncated to 27 bits that carry the action in their low mantissa bits, see pack_key)
//
// In eager PyTorch the same step is two batched GEMMs and nine elementwise / reduction passes over
// [2, 72, N] and [2, 18, N] tensors (5.4 GB of HBM traffic per 2 M envs, 1.45 ms); here the observations are
// read once (160 B per env) and two bytes per env are written.
//
// The contraction is tiny (K = 40, 72) and the kernel is bound by reading the observations, so it uses the
// warp-level tensor-core path (mma.sync m16n8k16, envs on the M axis) rather than a tcgen05 pipeline: a CTA
// stages a [48 features][128 envs] tile per agent in shared memory with 16-byte async copies (rows are
// env-contiguous = "MN-major"), ldmatrix.trans turns it into K-major A fragments, layer 1's accumulators are
// rectified, rounded to bf16 and reused in place as layer 2's A fragments — sixteen hidden units at a time, so
// that only two accumulator tiles of layer 1 are ever live (64 registers, 32 warps per SM: the kernel is a chain
// of dependent tensor-core, shared-memory and special-function latencies and needs the warps to hide them) —
// and the sample is an arg-max over the accumulator fragment plus two quad shuffles.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>

#include "../../include/pikazoo_b200.h"
#include "pz_policy.cuh"

namespace pzp {

constexpr int kThreads = 256;    // 8 warps x 16 envs
constexpr int kTileEnvs = 128;
#ifndef PZ_POLICY_MIN_CTAS
#define PZ_POLICY_MIN_CTAS 4
#endif
constexpr int kMinCtas = PZ_POLICY_MIN_CTAS;  // 4: 64 registers (44 B of spill loads), 32 warps per SM: 0.250 ms per 2 M envs; 3: 80 registers, 0.260 ms
constexpr int kKP = PZ_POLICY_MAX_FEATURES;  // 48: features padded to 3 k-steps of 16
constexpr int kHP = PZ_POLICY_MAX_HIDDEN;    // 80: hidden units padded to 10 n-tiles of 8 / 5 k-steps of 16
constexpr int kAP = PZ_POLICY_MAX_ACTIONS;   // 24: actions padded to 3 n-tiles of 8
// shared-memory row strides (elements) that make the fragment accesses conflict-free:
constexpr int kXS = kTileEnvs + 8;  // 272 B: the 8 rows of an ldmatrix 8x8 block start 4 banks apart
// The weights sit in shared memory in B-fragment order, so that one 128-bit load per lane fetches the operands of
// two tensor-core instructions:
//   w1f[agent][j][ks][lane] = {b0, b1 of hidden tile 2j, b0, b1 of hidden tile 2j + 1} for feature k-step ks
//   w2f[agent][j][grp][lane] = grp 0: {b0, b1 of action tile 0, b0, b1 of action tile 1}, grp 1: {b0, b1 of tile 2, 0, 0}
// with b0 = W[8 tile + g][16 kstep + 2t, +1], b1 = W[8 tile + g][16 kstep + 2t + 8, +9], lane = 4 g + t.
constexpr int kW1Words = 2 * (kHP / 16) * (kKP / 16) * 32 * 4;
constexpr int kW2Words = 2 * (kHP / 16) * 2 * 32 * 4;
static_assert(kAP == 24 && kHP % 16 == 0 && kKP % 16 == 0, "fragment layout");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t relu_pack(float lo, float hi) { return relu_pack_bf16x2(lo, hi); }

constexpr size_t kSmemBytes = sizeof(__nv_bfloat16) * 2 * kKP * kXS + sizeof(uint32_t) * (kW1Words + kW2Words);

// NA: the action count when it is one of the two the simulator has (18, or 13 under SimplifyAction), so that the
// padded candidates fall away at compile time; 0 = read it from the parameters.
template <int NA>
__global__ void __launch_bounds__(kThreads, kMinCtas) pz_policy_mlp_kernel(const __grid_constant__ Params P) {
    const int n_actions = NA ? NA : P.n_actions;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16(*xs)[kKP][kXS] = reinterpret_cast<__nv_bfloat16(*)[kKP][kXS]>(smem_raw);
    uint32_t *w1f = reinterpret_cast<uint32_t *>(xs + 2);
    uint32_t *w2f = w1f + kW1Words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // weights (zero-padded, fragment order) and the padding rows of the observation tile, once per CTA
    const __nv_bfloat16 zero = __float2bfloat16(0.0f);
    for (int i = tid; i < 2 * kW1Words; i += kThreads) {  // one bf16 each
        const int e = i & 1, idx = (i >> 1) & 3, ln = (i >> 3) & 31, blk = i >> 8;  // blk = (a * 5 + j) * 3 + ks
        const int ks = blk % (kKP / 16), j = (blk / (kKP / 16)) % (kHP / 16), a = blk / ((kKP / 16) * (kHP / 16));
        const int r = 16 * j + 8 * (idx >> 1) + (ln >> 2), c = 16 * ks + 8 * (idx & 1) + 2 * (ln & 3) + e;
        reinterpret_cast<__nv_bfloat16 *>(w1f)[i] =
            (r < P.h1 && c < P.k1) ? P.w1[((int64_t)a * P.h1 + r) * P.k1 + c] : zero;
    }
    for (int i = tid; i < 2 * kW2Words; i += kThreads) {
        const int e = i & 1, idx = (i >> 1) & 3, ln = (i >> 3) & 31, blk = i >> 8;  // blk = (a * 5 + j) * 2 + grp
        const int grp = blk & 1, j = (blk >> 1) % (kHP / 16), a = (blk >> 1) / (kHP / 16);
        const int nt = 2 * grp + (idx >> 1);  // action tile; tile 3 does not exist
        const int r = 8 * nt + (ln >> 2), c = 16 * j + 8 * (idx & 1) + 2 * (ln & 3) + e;
        reinterpret_cast<__nv_bfloat16 *>(w2f)[i] =
            (nt < kAP / 8 && r < n_actions && c < P.k2) ? P.w2[((int64_t)a * n_actions + r) * P.k2 + c] : zero;
    }
    for (int i = tid; i < 2 * kKP * kXS; i += kThreads) (&xs[0][0][0])[i] = zero;

    const bool vec_ok = (P.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(P.obs) & 15u) == 0);
    const int64_t n_tiles = (P.n + kTileEnvs - 1) / kTileEnvs;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t env0 = tile * kTileEnvs;
        __syncthreads();  // the previous tile has been consumed (and, first time, the fills above are done)
        if (vec_ok && env0 + kTileEnvs <= P.n) {
            // 16-byte pieces of 8 envs; a pass of the CTA covers kThreads / kPieces (agent, feature) rows. Row r of
            // the tile is global row r + (r >= k1) * (rows - k1) and shared row r + (r >= k1) * (kKP - k1).
            constexpr int kPieces = kTileEnvs / 8, kRowsPerPass = kThreads / kPieces;
            const int piece = tid % kPieces;
            const __nv_bfloat16 *src0 = P.obs + env0 + piece * 8;
            const uint32_t dst0 = smem_u32(&xs[0][0][piece * 8]);
            for (int row = tid / kPieces; row < 2 * P.k1; row += kRowsPerPass) {
                const int second = row >= P.k1 ? 1 : 0;
                const int grow = row + second * (P.rows - P.k1), srow = row + second * (kKP - P.k1);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(srow * kXS * 2)),
                             "l"(src0 + (int64_t)grow * P.ld)
                             : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {  // ragged last tile or unaligned rows
            for (int i = tid; i < 2 * P.k1 * kTileEnvs; i += kThreads) {
                const int a = i / (P.k1 * kTileEnvs), rem = i - a * (P.k1 * kTileEnvs), k = rem / kTileEnvs,
                          e = rem % kTileEnvs;
                xs[a][k][e] = env0 + e < P.n ? P.obs[((int64_t)a * P.rows + k) * P.ld + env0 + e] : zero;
            }
        }
        __syncthreads();

        const int m0 = warp * 16;
        const int64_t env_a = env0 + m0 + g, env_b = env_a + 8;  // the two rows this thread holds
        int act[2][2];                                           // [row half][agent]
        uint32_t nbase[2] = {0u, 0u};
        if (!P.greedy) {
            nbase[0] = noise_base(P.seed, P.step, P.first_env + (uint64_t)env_a);
            nbase[1] = noise_base(P.seed, P.step, P.first_env + (uint64_t)env_b);
        }
#pragma unroll
        for (int a = 0; a < 2; a++) {
            // layer 1, C1[env][hidden] = X^T[env][feature] . W1^T[feature][hidden], sixteen hidden units (two
            // accumulator tiles) at a time; rectified and rounded to bf16 that pair IS the A fragment of layer 2's
            // k-step j: C2[env][action] += relu(C1)[env][16 j ..] . W2^T[16 j ..][action]
            uint32_t xa[kKP / 16][4];
#pragma unroll
            for (int ks = 0; ks < kKP / 16; ks++) {
                const int j = lane >> 3, r = lane & 7;  // 8x8 block j: env half j & 1, feature half j >> 1
                const uint32_t addr = smem_u32(&xs[a][16 * ks + (j >> 1) * 8 + r][m0 + (j & 1) * 8]);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(xa[ks][0]), "=r"(xa[ks][1]), "=r"(xa[ks][2]), "=r"(xa[ks][3])
                             : "r"(addr));
            }
            float c2[kAP / 8][4];
#pragma unroll
            for (int nt = 0; nt < kAP / 8; nt++) c2[nt][0] = c2[nt][1] = c2[nt][2] = c2[nt][3] = 0.0f;
#pragma unroll
            for (int j = 0; j < kHP / 16; j++) {
                float ca[4] = {0.0f, 0.0f, 0.0f, 0.0f}, cb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int ks = 0; ks < kKP / 16; ks++) {
                    const uint4 b = reinterpret_cast<const uint4 *>(w1f)[((a * (kHP / 16) + j) * (kKP / 16) + ks) * 32 + lane];
                    mma_bf16(ca, xa[ks], b.x, b.y);
                    mma_bf16(cb, xa[ks], b.z, b.w);
                }
                uint32_t af[4];
                af[0] = relu_pack(ca[0], ca[1]);
                af[1] = relu_pack(ca[2], ca[3]);
                af[2] = relu_pack(cb[0], cb[1]);
                af[3] = relu_pack(cb[2], cb[3]);
                const uint4 b01 = reinterpret_cast<const uint4 *>(w2f)[((a * (kHP / 16) + j) * 2 + 0) * 32 + lane];
                const uint2 b2 = reinterpret_cast<const uint2 *>(w2f)[(((a * (kHP / 16) + j) * 2 + 1) * 32 + lane) * 2];
                mma_bf16(c2[0], af, b01.x, b01.y);
                mma_bf16(c2[1], af, b01.z, b01.w);
                mma_bf16(c2[2], af, b2.x, b2.y);
            }
            // this thread holds actions 8 nt + 2 t + {0, 1} of rows env_a (c[0], c[1]) and env_b (c[2], c[3])
            if (P.logits != nullptr) {  // launch-uniform
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int64_t env = h ? env_b : env_a;
#pragma unroll
                    for (int nt = 0; nt < kAP / 8; nt++)
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int action = 8 * nt + 2 * t + q;
                            if (action < n_actions && env < P.n)
                                P.logits[(env * 2 + a) * n_actions + action] = c2[nt][2 * h + q];
                        }
                }
            }
            // sample
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float best = pack_key(-INFINITY, 31);
#pragma unroll
                for (int nt = 0; nt < kAP / 8; nt++)
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int action = 8 * nt + 2 * t + q;
                        const float logit = c2[nt][2 * h + q];
                        if (action < n_actions)
                            best = fmaxf(best, pack_key(P.greedy ? logit : gumbel_key(logit, nbase[h], a, action), action));
                    }
                best = fmaxf(best, __shfl_xor_sync(0xFFFFFFFFu, best, 1));  // over the quad (t = 0..3)
                best = fmaxf(best, __shfl_xor_sync(0xFFFFFFFFu, best, 2));
                const int chosen = 31 - (int)(__float_as_uint(best) & 31u);
                act[h][a] = chosen < n_actions ? chosen : 0;  // every key NaN: action 0
            }
        }
        if (t == 0) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t env = h ? env_b : env_a;
                if (env >= P.n) continue;
                if (P.act_dtype == PZ_ACT_U8)
                    reinterpret_cast<uchar2 *>(P.actions)[env] = make_uchar2((unsigned char)act[h][0], (unsigned char)act[h][1]);
                else if (P.act_dtype == PZ_ACT_I32)
                    reinterpret_cast<int2 *>(P.actions)[env] = make_int2(act[h][0], act[h][1]);
                else
                    reinterpret_cast<longlong2 *>(P.actions)[env] = make_longlong2(act[h][0], act[h][1]);
            }
        }
    }
}

// Which kernel pz_policy_mlp_act launches (process-wide; the two produce the same sample definition)
static std::atomic<int> g_policy_impl{PZ_POLICY_IMPL_DEFAULT};

}  // namespace pzp

extern "C" int pz_policy_select(int32_t impl) {
    if (impl != PZ_POLICY_IMPL_TCGEN05 && impl != PZ_POLICY_IMPL_MMA_SYNC) return -1;
    return pzp::g_policy_impl.exchange(impl);
}

extern "C" int pz_policy_mlp_act(const void *obs_dev, int64_t n, int64_t ld, int32_t rows, const void *w1_dev,
                                 int32_t hidden_rows, int32_t features, const void *w2_dev, int32_t n_actions,
                                 int32_t w2_cols, uint64_t seed, uint64_t step, uint64_t first_env, void *actions_dev,
                                 int32_t action_dtype, int32_t greedy, float *logits_dev, void *stream) {
    using namespace pzp;
    if (!obs_dev || !w1_dev || !w2_dev || !actions_dev || n < 0 || ld < n) return PZ_E_BADARG;
    if (features < 1 || features > kKP || rows < features || hidden_rows < 1 || hidden_rows > kHP || w2_cols < 1 ||
        w2_cols > kHP || n_actions < 1 || n_actions > kAP)
        return PZ_E_BADCONFIG;
    if (action_dtype < PZ_ACT_I32 || action_dtype > PZ_ACT_U8) return PZ_E_BADCONFIG;
    if (n == 0) return 0;
    Params P;
    P.obs = reinterpret_cast<const __nv_bfloat16 *>(obs_dev);
    P.n = n;
    P.ld = ld;
    P.rows = rows;
    P.w1 = reinterpret_cast<const __nv_bfloat16 *>(w1_dev);
    P.w2 = reinterpret_cast<const __nv_bfloat16 *>(w2_dev);
    P.h1 = hidden_rows;
    P.k1 = features;
    P.n_actions = n_actions;
    P.k2 = w2_cols;
    P.seed = seed;
    P.step = step;
    P.first_env = first_env;
    P.actions = actions_dev;
    P.act_dtype = action_dtype;
    P.greedy = greedy != 0;
    P.logits = logits_dev;
    if (g_policy_impl.load(std::memory_order_relaxed) == PZ_POLICY_IMPL_TCGEN05) return launch_tc(P, (cudaStream_t)stream);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + kTileEnvs - 1) / kTileEnvs;
    const int64_t resident = (int64_t)sms * kMinCtas;  // persistent over tiles: the weights are staged once per CTA
    const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
    {  // more than 48 KB of dynamic shared memory needs the opt-in, once per device
        static std::mutex mu;
        static bool attr_set[64] = {};
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(pz_policy_mlp_kernel<18>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)kSmemBytes);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(pz_policy_mlp_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(pz_policy_mlp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
            if (e != cudaSuccess) return (int)e;
            attr_set[dev] = true;
        }
    }
    if (n_actions == 18)
        pz_policy_mlp_kernel<18><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    else if (n_actions == 13)
        pz_policy_mlp_kernel<13><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    else
        pz_policy_mlp_kernel<0><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(P);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : (int)err;
}
