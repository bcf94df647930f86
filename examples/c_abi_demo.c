/* The C ABI on its own: no Python, no torch. Builds with
 *     gcc examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include -Lpikazoo_b200/csrc -lpikazoo_b200 \
 *         -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/pikazoo_b200/csrc -o c_abi_demo
 * and prints two lines that tests/test_gpu_c_abi_demo.py compares with the oracle:
 *   device path: n envs, computer vs computer (actions_dev = NULL), T pz_step calls, then the unpacked
 *                state of every env is folded into one 64-bit checksum;
 *   host path:   pz_host_* with host buffers and the product's counter-based action stream, checksum over
 *                every observation / reward / done the T calls returned.
 * usage: c_abi_demo <n_envs> <steps> <seed>                                                         */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pikazoo_b200.h"

#define CHECK(x)                                                                       \
    do {                                                                               \
        int rc_ = (int)(x);                                                            \
        if (rc_ != 0) {                                                                \
            fprintf(stderr, "%s failed: %s (%d)\n", #x, pz_strerror(rc_), rc_);        \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

/* position-weighted byte sum folded into a running hash (mod 2^64); cheap to restate with numpy */
static uint64_t fold(uint64_t h, const void *data, size_t bytes) {
    const unsigned char *p = (const unsigned char *)data;
    uint64_t s = 0;
    for (size_t i = 0; i < bytes; i++) s += (uint64_t)p[i] * (uint64_t)(i + 1);
    return h * 0x100000001B3ULL + s;
}

/* the product's synthetic action stream, same formula as the rollout kernel's */
static int32_t synth_action(uint64_t seed, uint64_t env, uint64_t frame, int agent, uint32_t n_actions) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (2 * env + (uint64_t)agent + 1);
    z ^= frame * 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int32_t)(((z >> 32) * (uint64_t)n_actions) >> 32);
}

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 4096;
    const int steps = argc > 2 ? atoi(argv[2]) : 200;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;

    /* ---- device-pointer entry points: computer vs computer ---- */
    pz_config cfg;
    if (pz_config_init(&cfg, sizeof cfg)) return 3;
    cfg.is_player1_computer = cfg.is_player2_computer = 1;
    cfg.winning_score = 3;
    cfg.serve = PZ_SERVE_RANDOM;
    int32_t *state = NULL, *obs = NULL, *unpacked = NULL;
    int64_t *stats = NULL;
    CHECK(cudaMalloc((void **)&state, pz_state_bytes(n)));
    CHECK(cudaMalloc((void **)&obs, (size_t)n * 2 * PZ_OBS_WORDS * sizeof(int32_t)));
    CHECK(cudaMalloc((void **)&unpacked, (size_t)n * PZ_UNPACKED_WORDS * sizeof(int32_t)));
    CHECK(cudaMalloc((void **)&stats, PZ_NUM_STATS * sizeof(int64_t)));
    CHECK(cudaMemset(stats, 0, PZ_NUM_STATS * sizeof(int64_t)));
    CHECK(pz_seed(state, n, seed, 0, NULL));
    CHECK(pz_reset(state, n, &cfg, obs, NULL));
    for (int t = 0; t < steps; t++) CHECK(pz_step(state, n, &cfg, NULL, obs, NULL, NULL, stats, NULL));
    CHECK(pz_export_state(state, n, unpacked, NULL));
    int32_t *h_unpacked = (int32_t *)malloc((size_t)n * PZ_UNPACKED_WORDS * sizeof(int32_t));
    int64_t h_stats[PZ_NUM_STATS];
    CHECK(cudaMemcpy(h_unpacked, unpacked, (size_t)n * PZ_UNPACKED_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost));
    printf("device %lld %d %016llx episodes=%lld resets=%lld\n", (long long)n, steps,
           (unsigned long long)fold(0xCBF29CE484222325ULL, h_unpacked, (size_t)n * PZ_UNPACKED_WORDS * 4),
           (long long)h_stats[PZ_STAT_EPISODES], (long long)h_stats[PZ_STAT_RESETS]);

    /* ---- host-buffer entry points: random actions, fused SimplifyAction ---- */
    pz_config hc;
    if (pz_config_init(&hc, sizeof hc)) return 3;
    hc.simplify_action = 1;
    hc.winning_score = 2;
    pz_host_ctx *ctx = NULL;
    CHECK(pz_host_create(&ctx, n, &hc, seed, 0, 4));
    int32_t *h_act = (int32_t *)malloc((size_t)n * 2 * sizeof(int32_t));
    int32_t *h_obs = (int32_t *)malloc((size_t)n * 2 * PZ_OBS_WORDS * sizeof(int32_t));
    float *h_rew = (float *)malloc((size_t)n * 2 * sizeof(float));
    uint8_t *h_done = (uint8_t *)malloc((size_t)n);
    CHECK(pz_host_reset(ctx, h_obs));
    uint64_t h = fold(0xCBF29CE484222325ULL, h_obs, (size_t)n * 2 * PZ_OBS_WORDS * 4);
    for (int t = 0; t < steps; t++) {
        for (int64_t i = 0; i < n; i++)
            for (int a = 0; a < 2; a++) h_act[2 * i + a] = synth_action(7, (uint64_t)i, (uint64_t)t, a, 13);
        CHECK(pz_host_step(ctx, h_act, h_obs, h_rew, h_done));
        h = fold(h, h_obs, (size_t)n * 2 * PZ_OBS_WORDS * 4);
        h = fold(h, h_rew, (size_t)n * 2 * sizeof(float));
        h = fold(h, h_done, (size_t)n);
    }
    printf("host %lld %d %016llx\n", (long long)n, steps, (unsigned long long)h);
    pz_host_destroy(ctx);
    cudaFree(state), cudaFree(obs), cudaFree(unpacked), cudaFree(stats);
    free(h_unpacked), free(h_act), free(h_obs), free(h_rew), free(h_done);
    return 0;
}
