// Host side of the compact wire format of the host-buffer path (pz_host_set_wire, pz_wire_expand): the link, not the
// kernel, bounds a caller behind PCIe, and what crosses it is redundant — player_2's observation
// is player_1's with the two 13-value player blocks swapped (pikazoo/env/pikazoo_env.py:585-586), every value fits
// int16, and the rewards and the termination flag of an unshaped step (pikazoo_env.py:190-240) are three values in
// one byte. So the device ships player_1's int16 row + a status byte (71 B per env instead of 289 B) and THIS file
// rebuilds, on host threads, exactly the arrays the reference's caller expects: obs [n][2][35], reward [n][2], done
// [n]. Pure host code (no CUDA): compiled by the host compiler, AVX2 path selected at run time.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/pikazoo_b200.h"

namespace pzw {

constexpr int kRow = PZ_OBS_WORDS;  // 35

// ---- portable form ---------------------------------------------------------------------------------------------
template <typename T>
static inline void expand_row(const int16_t *s, T *d) {
    for (int j = 0; j < kRow; j++) d[j] = (T)s[j];                   // player_1: the row as it is
    T *q = d + kRow;                                                 // player_2 (pikazoo_env.py:586):
    for (int j = 0; j < 13; j++) q[j] = (T)s[13 + j];                //   its own block first,
    for (int j = 0; j < 13; j++) q[13 + j] = (T)s[j];                //   then the opponent's,
    for (int j = 26; j < kRow; j++) q[j] = (T)s[j];                  //   then the ball
}

template <typename T>
static void expand_obs_scalar(const int16_t *rows, int64_t n, T *obs) {
    for (int64_t i = 0; i < n; i++) expand_row<T>(rows + i * kRow, obs + i * 2 * kRow);
}

#if defined(__x86_64__)
// ---- AVX2, int32 output: 8 envs = 560 int32 = 70 aligned 32-byte vectors, built in an L1-resident block with
// overlapping 8-wide sign-extending copies (35 = 8 + 8 + 8 + 8 + 3: the last copy of a run starts 3 early), then
// written with non-temporal stores: the output (280 B per env) is never read by this thread again, and a
// write-allocate would double the memory traffic of what is a memory-bound pass.
__attribute__((target("avx2"))) static inline void cvt8(const int16_t *s, int32_t *d) {
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(d),
                        _mm256_cvtepi16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(s))));
}
__attribute__((target("avx2"))) static void expand_obs_i32_avx2(const int16_t *rows, int64_t n, int32_t *obs) {
    alignas(32) int32_t blk[8 * 2 * kRow];
    int64_t i = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(obs) & 31u) == 0;
    // the overlapping 8-wide reads stay inside the row being converted, so the last row needs no special case
    for (; i + 8 <= n; i += 8) {
        for (int e = 0; e < 8; e++) {
            const int16_t *s = rows + (i + e) * kRow;
            int32_t *d = blk + e * 2 * kRow;
            cvt8(s, d), cvt8(s + 8, d + 8), cvt8(s + 16, d + 16), cvt8(s + 24, d + 24), cvt8(s + 27, d + 27);
            int32_t *q = d + kRow;
            cvt8(s + 13, q), cvt8(s + 18, q + 5);        // own block   s[13..26) -> q[0..13)
            cvt8(s, q + 13), cvt8(s + 5, q + 18);        // opponent    s[0..13)  -> q[13..26)
            cvt8(s + 26, q + 26), cvt8(s + 27, q + 27);  // ball        s[26..35) -> q[26..35)
        }
        __m256i *out = reinterpret_cast<__m256i *>(obs + i * 2 * kRow);
        const __m256i *in = reinterpret_cast<const __m256i *>(blk);
        if (aligned) {
            for (int v = 0; v < 70; v++) _mm256_stream_si256(out + v, _mm256_load_si256(in + v));
        } else {
            for (int v = 0; v < 70; v++) _mm256_storeu_si256(out + v, _mm256_load_si256(in + v));
        }
    }
    if (aligned) _mm_sfence();
    expand_obs_scalar<int32_t>(rows + i * kRow, n - i, obs + i * 2 * kRow);
}
static bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#endif

template <typename R>
static void expand_status(const uint8_t *status, int64_t n, R *reward, uint8_t *done) {
    if (reward)
        for (int64_t i = 0; i < n; i++) {  // status: (player_1's reward + 1) | done << 2 | truncated << 3
            const int base = (int)(status[i] & 3u) - 1;
            reward[2 * i] = (R)base;        // pikazoo_env.py:200-201 / :211-212: +1 / -1 for the scorer, the negative
            reward[2 * i + 1] = (R)(-base);  // for the other; 0 / -0 do not occur: a frame without a point is 0, 0
        }
    if (done)
        for (int64_t i = 0; i < n; i++) done[i] = (uint8_t)((status[i] >> 2) & 1u);
}

}  // namespace pzw

extern "C" int pz_wire_expand(const int16_t *rows, const uint8_t *status, int64_t n, int32_t obs_dtype, void *obs,
                              int32_t reward_dtype, void *reward, uint8_t *done) {
    using namespace pzw;
    if (n < 0 || (obs && !rows) || ((reward || done) && !status)) return PZ_E_BADARG;
    if (obs) {
        if (obs_dtype == PZ_OBS_I32) {
#if defined(__x86_64__)
            if (have_avx2())
                expand_obs_i32_avx2(rows, n, static_cast<int32_t *>(obs));
            else
#endif
                expand_obs_scalar<int32_t>(rows, n, static_cast<int32_t *>(obs));
        } else if (obs_dtype == PZ_OBS_I16) {
            expand_obs_scalar<int16_t>(rows, n, static_cast<int16_t *>(obs));
        } else {
            return PZ_E_BADCONFIG;  // normalised / floating-point rows are not a wire format
        }
    }
    if (reward_dtype == PZ_REW_F32)
        expand_status<float>(status, n, static_cast<float *>(reward), done);
    else if (reward_dtype == PZ_REW_F64)
        expand_status<double>(status, n, static_cast<double *>(reward), done);
    else if (reward || done)
        return PZ_E_BADCONFIG;
    return 0;
}
