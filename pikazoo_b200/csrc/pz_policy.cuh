// Shared by the two implementations of pz_policy_mlp_act (pz_policy.cu: warp-level mma.sync; pz_policy_tc.cu:
// tcgen05 + TMEM): the launch parameters and the counter-based Gumbel noise / packed-key arg-max, which are the
// definition of the sample and therefore identical in both.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pikazoo_b200.h"

namespace pzp {

struct Params {
    const __nv_bfloat16 *obs;  // [2][rows][ld], element (a, k, env)
    int64_t n, ld;
    int rows;
    const __nv_bfloat16 *w1;  // [2][h1][k1]
    const __nv_bfloat16 *w2;  // [2][n_actions][k2]
    int h1, k1, n_actions, k2;
    uint64_t seed, step, first_env;
    void *actions;
    int act_dtype, greedy;
    float *logits;  // optional [n][2][n_actions]
};

#ifdef PZ_HOST_EMULATION  // tests/emul: this header compiled for the host; libm stands in for the MUFU approximations
__device__ __forceinline__ float lg2_approx(float x) { return log2f(x); }
__device__ __forceinline__ float ex2_approx(float x) { return exp2f(x); }
#else
__device__ __forceinline__ float lg2_approx(float x) {  // MUFU.LG2; the arguments here are normal numbers
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Two fp32 values rectified and rounded to a packed bf16 pair (low half = lo) in one conversion instruction
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
#endif

// Counter-based noise, restated in pikazoo_b200/policy.py (gumbel_noise_reference) for the tests: one 64-bit
// mix per (seed, step, global env), one 32-bit mix per (agent, action). The key added to a logit is
//   key = fma(-ln2, log2(-log2(u)), logit) = logit + Gumbel(u) + ln(ln 2): the constant does not move the arg-max.
__device__ __forceinline__ uint32_t noise_base(uint64_t seed, uint64_t step, uint64_t genv) {
    uint64_t z = (seed + 0x9E3779B97F4A7C15ULL * (genv + 1ULL)) ^ (step * 0xD1B54A32D192ED03ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (uint32_t)z;
}
// u = ((x >> 9) + 0.5) / 2^23 in (0, 1), exact: the 23 bits are placed as the mantissa of a float in [1, 2) and
// 1 - 2^-24 is subtracted (both steps exact) — one logic and one add instruction instead of a conversion on the
// special-function pipe, which the two logarithms already load
// 32 mixed bits -> u = ((x >> 9) + 0.5) / 2^23 in (0, 1)
__device__ __forceinline__ float uniform_from_counter(uint32_t x) {
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return __uint_as_float(0x3F800000u | (x >> 9)) + (-1.0f + 5.9604644775390625e-08f);
}
__device__ __forceinline__ float gumbel_log_term(uint32_t x) {  // log2(-log2(u)): everything but the logit
    return lg2_approx(-lg2_approx(uniform_from_counter(x)));
}
__device__ __forceinline__ float gumbel_from_bits(float logit, uint32_t x) {
    return fmaf(-0.693147182f, gumbel_log_term(x), logit);
}
__device__ __forceinline__ float gumbel_key(float logit, uint32_t base, int agent, int action) {
    return gumbel_from_bits(logit, base + (uint32_t)(32 * agent + action + 1) * 0x9E3779B9u);  // base is already well mixed
}
// the same with `agent_base` = base + 32 * agent * 0x9E3779B9 hoisted by the caller
__device__ __forceinline__ float gumbel_key_from(float logit, uint32_t agent_base, int action) {
    return gumbel_from_bits(logit, agent_base + (uint32_t)(action + 1) * 0x9E3779B9u);
}

// The arg-max runs on keys that carry their action in the five low mantissa bits (31 - action, so that among
// positive keys equal in the upper 27 bits the lower action wins): one LOP3 + one FMNMX per candidate and one
// shuffle + one FMNMX per reduction step instead of compare-and-select pairs on (key, action). The 2^-19
// relative truncation of the key is far below the noise resolution.
__device__ __forceinline__ float pack_key(float key, int action) {
    return __uint_as_float((__float_as_uint(key) | 31u) ^ (uint32_t)action);  // low five bits = 31 - action: one LOP3
}

// The tcgen05 kernel's categorical sample (one thread holds all the logits of its env and agent): inversion of the
// cumulative distribution with ONE uniform per (env, agent) — the first uniform of the Gumbel stream above, counter
// (seed, step, global env, agent, action 0):
//   w_j = 2^((logit_j - max) * log2 e),  c_j = w_0 + ... + w_j (fp32, in action order),
//   action = #{ j < n_actions - 1 : c_j <= u * c_last }
// i.e. P(action = j) = softmax(logits)_j up to the 2^-23 resolution of u. Restated in pikazoo_b200/policy.py
// (inverse_cdf_reference). A third of the Gumbel arg-max's instructions and half of its special-function work.
template <int N>
__device__ __forceinline__ int sample_inverse_cdf(const float (&logit)[N], int n_actions, uint32_t agent_base) {
    float m = logit[0];
#pragma unroll
    for (int j = 1; j < N; j++)
        if (j < n_actions) m = fmaxf(m, logit[j]);
    const float shift = -m * 1.44269504f;
    float c[N], acc = 0.0f;
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (j < n_actions) acc += ex2_approx(fmaf(logit[j], 1.44269504f, shift));
        c[j] = acc;
    }
    const float target = uniform_from_counter(agent_base + 0x9E3779B9u) * acc;
    int action = 0;
#ifdef PZ_HOST_EMULATION
#pragma unroll
    for (int j = 0; j < N - 1; j++)
        if (j < n_actions - 1) action += (c[j] <= target) ? 1 : 0;
#else
    // one FSET (all-ones when true) per boundary and one three-input add per two of them, instead of a
    // compare / select / add triple each
#pragma unroll
    for (int j = 0; j < N - 1; j++)
        if (j < n_actions - 1) {
            int m;
            asm("set.le.s32.f32 %0, %1, %2;" : "=r"(m) : "f"(c[j]), "f"(target));
            action -= m;
        }
#endif
    return action;  // NaN logits: every comparison is false, action 0
}

// pz_policy_tc.cu
int launch_tc(const Params &P, cudaStream_t stream);

}  // namespace pzp
