// Host-buffer path of the C ABI (pz_host_*): what a numpy caller of the reference binds.
// One context owns the device state and staging buffers of n envs; a step is split into
// `chunks` env ranges, each on its own stream: H2D(actions) -> pz_step_kernel -> D2H(obs,
// reward, done, status), so the PCIe copies of one chunk overlap the kernel and copies of the others.
// pz_host_step_begin enqueues all of it and returns; pz_host_step_end waits. The link is the bound
// (DESIGN.md section 5), so the bytes are what matters: the shared-row int16 observation layout plus the one-byte
// status (reward, done, truncated) is 71 B per env-step against 297 B for the reference's dtypes.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "pz_kernels.cuh"

// Compact wire format (pz_host_set_wire): the step kernel writes player_1's int16 row and the status byte, those
// 71 B per env cross the link into a pinned staging buffer owned by the context, and a pool of host threads expands
// each chunk, as soon as its copy has landed, into the caller's arrays (pz_wire_expand, pz_wire.cpp) while the
// following chunks are still on the link. The caller sees the reference's dtypes and layout; PCIe sees a quarter
// of the bytes.
struct pz_wire_pool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t generation = 0;  // bumped per job
    int remaining = 0;        // workers still busy with the current job
    bool quit = false;
    std::atomic<int> error{0};
    // the current job
    void *obs = nullptr, *reward = nullptr;
    uint8_t *done = nullptr, *status_out = nullptr;
};

struct pz_host_ctx {
    int64_t n = 0;
    pz_config cfg{};
    int device = 0;
    int32_t *state = nullptr;
    void *actions = nullptr;
    char *obs = nullptr;
    void *reward = nullptr;
    uint8_t *done = nullptr;
    uint8_t *status = nullptr;
    int64_t *stats = nullptr;
    std::vector<cudaStream_t> streams;
    std::vector<int64_t> bounds;  // chunk c = [bounds[c], bounds[c+1])
    size_t act_elem = 4, rew_elem = 4, obs_row = 2 * PZ_OBS_WORDS * 4;  // obs_row: bytes per env
    bool in_flight = false;
    // compact wire format
    int wire = PZ_WIRE_NATIVE;
    bool wire_rewards = false;        // rewards and done are rebuilt from the status byte (unshaped rewards)
    pz_config wire_cfg{};             // cfg with int16 shared rows
    int16_t *wire_obs_dev = nullptr;  // [n][35]
    int16_t *wire_obs_host = nullptr; // pinned
    uint8_t *wire_status_host = nullptr;
    std::vector<cudaEvent_t> events;  // chunk k's copies have landed
    pz_wire_pool *pool = nullptr;
};

namespace {

#define PZ_CUDA(x)                             \
    do {                                       \
        cudaError_t e_ = (x);                  \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

// Every entry point runs on the context's device whatever device is current in the calling thread.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

int sync_all(pz_host_ctx *c) {
    for (cudaStream_t s : c->streams) PZ_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// One worker of the pool: for every job, walk the chunks in the order their copies were enqueued, wait for chunk
// k's event and expand slice `rank` of `count` of it. All workers share every chunk, so the expansion of chunk k runs
// on all threads while chunk k + 1 is on the link.
void wire_worker(pz_host_ctx *c, int rank, int count) {
    pz_wire_pool *p = c->pool;
    cudaSetDevice(c->device);
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(p->mu);
            p->cv_go.wait(lk, [&] { return p->quit || p->generation != seen; });
            if (p->quit) return;
            seen = p->generation;
        }
        const int chunks = (int)c->streams.size();
        for (int k = 0; k < chunks; k++) {
            const int64_t b = c->bounds[k], e = c->bounds[k + 1];
            if (e <= b) continue;
            const cudaError_t err = cudaEventSynchronize(c->events[k]);
            if (err != cudaSuccess) {
                p->error.store((int)err);
                break;
            }
            // slices on multiples of 8 envs: 32-byte aligned pieces of every output array
            const int64_t blocks = (e - b + 7) / 8;
            int64_t lo = b + (blocks * rank / count) * 8, hi = b + (blocks * (rank + 1) / count) * 8;
            if (hi > e) hi = e;
            if (lo >= hi) continue;
            const bool st = p->status_out != nullptr || c->wire_rewards;
            const int rc = pz_wire_expand(
                p->obs ? c->wire_obs_host + lo * PZ_OBS_WORDS : nullptr, st ? c->wire_status_host + lo : nullptr, hi - lo,
                c->cfg.obs_dtype, p->obs ? (char *)p->obs + (size_t)lo * c->obs_row : nullptr, c->cfg.reward_dtype,
                c->wire_rewards && p->reward ? (char *)p->reward + (size_t)lo * 2 * c->rew_elem : nullptr,
                c->wire_rewards && p->done ? p->done + lo : nullptr);
            if (rc) p->error.store(rc);
            if (p->status_out) memcpy(p->status_out + lo, c->wire_status_host + lo, (size_t)(hi - lo));
        }
        {
            std::lock_guard<std::mutex> lk(p->mu);
            if (--p->remaining == 0) p->cv_done.notify_all();
        }
    }
}

void wire_start_job(pz_host_ctx *c, void *obs, void *reward, uint8_t *done, uint8_t *status_out) {
    pz_wire_pool *p = c->pool;
    std::lock_guard<std::mutex> lk(p->mu);
    p->obs = obs, p->reward = reward, p->done = done, p->status_out = status_out;
    p->remaining = (int)p->threads.size();
    p->generation++;
    p->cv_go.notify_all();
}

int wire_finish_job(pz_host_ctx *c) {
    pz_wire_pool *p = c->pool;
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->remaining == 0; });
    return p->error.exchange(0);
}

void wire_teardown(pz_host_ctx *c) {
    if (c->pool) {
        {
            std::lock_guard<std::mutex> lk(c->pool->mu);
            c->pool->quit = true;
            c->pool->cv_go.notify_all();
        }
        for (std::thread &t : c->pool->threads) t.join();
        delete c->pool;
        c->pool = nullptr;
    }
    for (cudaEvent_t e : c->events)
        if (e) cudaEventDestroy(e);
    c->events.clear();
    cudaFree(c->wire_obs_dev);
    cudaFreeHost(c->wire_obs_host);
    cudaFreeHost(c->wire_status_host);
    c->wire_obs_dev = nullptr, c->wire_obs_host = nullptr, c->wire_status_host = nullptr;
    c->wire = PZ_WIRE_NATIVE;
}

}  // namespace

extern "C" {

int pz_host_set_wire(pz_host_ctx *c, int32_t mode, int32_t threads) {
    if (!c || c->in_flight || (mode != PZ_WIRE_NATIVE && mode != PZ_WIRE_COMPACT)) return PZ_E_BADARG;
    DeviceGuard guard(c->device);
    if (int rc = sync_all(c)) return rc;
    wire_teardown(c);
    if (mode == PZ_WIRE_NATIVE) return 0;
    // what the wire carries is the integer observation in the reference's env-major rows
    if ((c->cfg.obs_dtype != PZ_OBS_I32 && c->cfg.obs_dtype != PZ_OBS_I16) || c->cfg.obs_layout != PZ_LAYOUT_ENV_MAJOR)
        return PZ_E_BADCONFIG;
    if (threads < 1) {
        const unsigned hw = std::thread::hardware_concurrency();
        threads = (int32_t)(hw == 0 ? 4 : (hw > 64 ? 64 : hw));
    }
    c->wire_cfg = c->cfg;
    c->wire_cfg.obs_dtype = PZ_OBS_I16;
    c->wire_cfg.obs_layout = PZ_LAYOUT_ENV_MAJOR_SHARED;
    // shaped rewards (RewardByBallPosition / RewardInNormalState) are not a function of the status byte: they travel
    c->wire_rewards = !c->cfg.reward_by_ball_position && c->cfg.reward_in_normal_state == PZ_RINS_OFF;
    const size_t n = (size_t)c->n;
    int rc = 0;
    do {
        if ((rc = (int)cudaMalloc(&c->wire_obs_dev, n * PZ_OBS_WORDS * sizeof(int16_t)))) break;
        if ((rc = (int)cudaHostAlloc(&c->wire_obs_host, n * PZ_OBS_WORDS * sizeof(int16_t), cudaHostAllocDefault))) break;
        if ((rc = (int)cudaHostAlloc(&c->wire_status_host, n, cudaHostAllocDefault))) break;
        c->events.assign(c->streams.size(), nullptr);
        for (size_t k = 0; k < c->events.size() && !rc; k++)
            rc = (int)cudaEventCreateWithFlags(&c->events[k], cudaEventDisableTiming);
        if (rc) break;
        c->pool = new (std::nothrow) pz_wire_pool();
        if (!c->pool) {
            rc = (int)cudaErrorMemoryAllocation;
            break;
        }
        for (int t = 0; t < threads; t++) c->pool->threads.emplace_back(wire_worker, c, t, (int)threads);
    } while (0);
    if (rc) {
        wire_teardown(c);
        return rc;
    }
    c->wire = PZ_WIRE_COMPACT;
    return 0;
}

int pz_host_create(pz_host_ctx **out, int64_t n, const pz_config *cfg, uint64_t base_seed, uint64_t first_env,
                   int32_t chunks) {
    if (!out || !cfg || n < 1) return PZ_E_BADARG;
    if (int rc = pz::check_config(cfg)) return rc;  // the ABI handshake comes first
    pz_host_ctx *c = new (std::nothrow) pz_host_ctx();
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->n = n;
    c->cfg = *cfg;
    c->act_elem = cfg->action_dtype == PZ_ACT_I64 ? 8 : (cfg->action_dtype == PZ_ACT_U8 ? 1 : 4);
    c->rew_elem = cfg->reward_dtype == PZ_REW_F64 ? 8 : 4;
    if (pz_obs_elem_bytes(cfg->obs_dtype) == 0 || cfg->obs_layout == PZ_LAYOUT_FEATURE_MAJOR) {  // host rows are per env
        delete c;
        return PZ_E_BADCONFIG;
    }
    c->obs_row = (cfg->obs_layout == PZ_LAYOUT_ENV_MAJOR_SHARED ? 1 : 2) * PZ_OBS_WORDS * pz_obs_elem_bytes(cfg->obs_dtype);
    if (chunks < 1) {  // automatic: one chunk per 128 Ki envs, at most 8 — small batches are bound by API calls, not PCIe
        const int64_t want = n / (128 * 1024);
        chunks = (int32_t)(want < 1 ? 1 : (want > 8 ? 8 : want));
    }
    // chunk boundaries on multiples of 128 envs (whole CTAs, 16-byte aligned slices of every array)
    int64_t blocks = (n + 127) / 128;
    if (chunks > blocks) chunks = (int32_t)blocks;
    c->bounds.resize(chunks + 1);
    for (int k = 0; k <= chunks; k++) {
        int64_t b = (blocks * k / chunks) * 128;
        c->bounds[k] = b > n ? n : b;
    }
    c->bounds[chunks] = n;
    int rc = 0;
    do {
        if ((rc = (int)cudaGetDevice(&c->device))) break;
        if ((rc = (int)cudaMalloc(&c->state, pz_state_bytes(n)))) break;
        if ((rc = (int)cudaMalloc(&c->actions, (size_t)n * 2 * c->act_elem))) break;
        if ((rc = (int)cudaMalloc(&c->obs, (size_t)n * c->obs_row))) break;
        if ((rc = (int)cudaMalloc(&c->reward, (size_t)n * 2 * c->rew_elem))) break;
        if ((rc = (int)cudaMalloc(&c->done, (size_t)n))) break;
        if ((rc = (int)cudaMalloc(&c->status, (size_t)n))) break;
        if ((rc = (int)cudaMalloc(&c->stats, PZ_NUM_STATS * sizeof(int64_t)))) break;
        c->streams.resize(chunks);
        for (int k = 0; k < chunks && !rc; k++)
            rc = (int)cudaStreamCreateWithFlags(&c->streams[k], cudaStreamNonBlocking);
        if (rc) break;
        // the statistics are zeroed and the envs seeded on stream 0; the device-wide synchronisation below orders
        // both before anything the (non-blocking) chunk streams will ever do
        if ((rc = (int)cudaMemsetAsync(c->stats, 0, PZ_NUM_STATS * sizeof(int64_t), c->streams[0]))) break;
        if ((rc = pz_seed(c->state, n, base_seed, first_env, c->streams[0]))) break;
        rc = (int)cudaDeviceSynchronize();
    } while (0);
    if (rc) {
        pz_host_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

int pz_host_reset(pz_host_ctx *c, void *obs_host) {
    if (!c || c->in_flight) return PZ_E_BADARG;
    DeviceGuard guard(c->device);
    const int chunks = (int)c->streams.size();
    if (c->wire == PZ_WIRE_COMPACT && obs_host) {
        for (int k = 0; k < chunks; k++) {
            const int64_t b = c->bounds[k], e = c->bounds[k + 1];
            if (e <= b) continue;
            int rc = pz::launch_reset(c->state, c->n, b, e, &c->wire_cfg, c->wire_obs_dev, nullptr, c->streams[k]);
            if (rc) return rc;
            PZ_CUDA(cudaMemcpyAsync(c->wire_obs_host + b * PZ_OBS_WORDS, c->wire_obs_dev + b * PZ_OBS_WORDS,
                                    (size_t)(e - b) * PZ_OBS_WORDS * sizeof(int16_t), cudaMemcpyDeviceToHost, c->streams[k]));
        }
        if (int rc = sync_all(c)) return rc;
        return pz_wire_expand(c->wire_obs_host, nullptr, c->n, c->cfg.obs_dtype, obs_host, c->cfg.reward_dtype, nullptr,
                              nullptr);
    }
    for (int k = 0; k < chunks; k++) {
        const int64_t b = c->bounds[k], e = c->bounds[k + 1];
        if (e <= b) continue;
        int rc = pz::launch_reset(c->state, c->n, b, e, &c->cfg, obs_host ? c->obs : nullptr, nullptr, c->streams[k]);
        if (rc) return rc;
        if (obs_host)
            PZ_CUDA(cudaMemcpyAsync((char *)obs_host + (size_t)b * c->obs_row, c->obs + (size_t)b * c->obs_row,
                                    (size_t)(e - b) * c->obs_row, cudaMemcpyDeviceToHost, c->streams[k]));
    }
    return sync_all(c);
}

// One chunk of a step with the arrays travelling as the caller will see them.
static int enqueue_chunk_native(pz_host_ctx *c, int k, const void *actions_host, void *obs_host, void *reward_host,
                                uint8_t *done_host, uint8_t *status_host) {
    const int64_t b = c->bounds[k], e = c->bounds[k + 1];
    if (e <= b) return 0;
    pz_episode_io ep;
    memset(&ep, 0, sizeof(ep));
    ep.status_dev = status_host ? c->status : nullptr;
    cudaStream_t s = c->streams[k];
    const size_t cnt = (size_t)(e - b);
    if (actions_host)
        PZ_CUDA(cudaMemcpyAsync((char *)c->actions + (size_t)b * 2 * c->act_elem,
                                (const char *)actions_host + (size_t)b * 2 * c->act_elem, cnt * 2 * c->act_elem,
                                cudaMemcpyHostToDevice, s));
    int rc = pz::launch_step(c->state, c->n, b, e, &c->cfg, actions_host ? c->actions : nullptr,
                             obs_host ? c->obs : nullptr, reward_host ? c->reward : nullptr,
                             done_host ? c->done : nullptr, c->stats, status_host ? &ep : nullptr, s);
    if (rc) return rc;
    if (obs_host)
        PZ_CUDA(cudaMemcpyAsync((char *)obs_host + (size_t)b * c->obs_row, c->obs + (size_t)b * c->obs_row,
                                cnt * c->obs_row, cudaMemcpyDeviceToHost, s));
    if (reward_host)
        PZ_CUDA(cudaMemcpyAsync((char *)reward_host + (size_t)b * 2 * c->rew_elem,
                                (const char *)c->reward + (size_t)b * 2 * c->rew_elem, cnt * 2 * c->rew_elem,
                                cudaMemcpyDeviceToHost, s));
    if (done_host) PZ_CUDA(cudaMemcpyAsync(done_host + b, c->done + b, cnt, cudaMemcpyDeviceToHost, s));
    if (status_host) PZ_CUDA(cudaMemcpyAsync(status_host + b, c->status + b, cnt, cudaMemcpyDeviceToHost, s));
    return 0;
}

// One chunk of a step in the compact wire format: the int16 row and the status byte land in the context's pinned
// staging buffers; the chunk's event tells the workers when.
static int enqueue_chunk_compact(pz_host_ctx *c, int k, const void *actions_host, void *obs_host, void *reward_host,
                                 uint8_t *done_host, uint8_t *status_host) {
    const int64_t b = c->bounds[k], e = c->bounds[k + 1];
    if (e <= b) return 0;
    const bool need_status = status_host || (c->wire_rewards && (reward_host || done_host));
    const bool native_rewards = !c->wire_rewards;
    pz_episode_io ep;
    memset(&ep, 0, sizeof(ep));
    ep.status_dev = need_status ? c->status : nullptr;
    cudaStream_t s = c->streams[k];
    const size_t cnt = (size_t)(e - b);
    if (actions_host)
        PZ_CUDA(cudaMemcpyAsync((char *)c->actions + (size_t)b * 2 * c->act_elem,
                                (const char *)actions_host + (size_t)b * 2 * c->act_elem, cnt * 2 * c->act_elem,
                                cudaMemcpyHostToDevice, s));
    int rc = pz::launch_step(c->state, c->n, b, e, &c->wire_cfg, actions_host ? c->actions : nullptr,
                             obs_host ? c->wire_obs_dev : nullptr, native_rewards && reward_host ? c->reward : nullptr,
                             native_rewards && done_host ? c->done : nullptr, c->stats, need_status ? &ep : nullptr, s);
    if (rc) return rc;
    if (obs_host)
        PZ_CUDA(cudaMemcpyAsync(c->wire_obs_host + b * PZ_OBS_WORDS, c->wire_obs_dev + b * PZ_OBS_WORDS,
                                cnt * PZ_OBS_WORDS * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
    if (need_status) PZ_CUDA(cudaMemcpyAsync(c->wire_status_host + b, c->status + b, cnt, cudaMemcpyDeviceToHost, s));
    if (native_rewards && reward_host)
        PZ_CUDA(cudaMemcpyAsync((char *)reward_host + (size_t)b * 2 * c->rew_elem,
                                (const char *)c->reward + (size_t)b * 2 * c->rew_elem, cnt * 2 * c->rew_elem,
                                cudaMemcpyDeviceToHost, s));
    if (native_rewards && done_host) PZ_CUDA(cudaMemcpyAsync(done_host + b, c->done + b, cnt, cudaMemcpyDeviceToHost, s));
    PZ_CUDA(cudaEventRecord(c->events[k], s));
    return 0;
}

int pz_host_step_begin(pz_host_ctx *c, const void *actions_host, void *obs_host, void *reward_host,
                       uint8_t *done_host, uint8_t *status_host) {
    if (!c || c->in_flight) return PZ_E_BADARG;
    const bool both_ai = c->cfg.is_player1_computer && c->cfg.is_player2_computer;
    if (!actions_host && !both_ai) return PZ_E_BADARG;
    DeviceGuard guard(c->device);
    const int chunks = (int)c->streams.size();
    // (A hybrid — the last chunks travelling natively behind the compact ones, so that the DMA engine and the host
    // threads deliver at the same time — was measured and dropped: 330 M env-steps/s either way on the 16-core B200
    // boxes, profiles/r02_host_wire_sweep.json; DMA writes and the threads' stores share the host's memory write
    // bandwidth, which is what bounds the compact format there.)
    for (int k = 0; k < chunks; k++) {
        const int rc = c->wire == PZ_WIRE_COMPACT
                           ? enqueue_chunk_compact(c, k, actions_host, obs_host, reward_host, done_host, status_host)
                           : enqueue_chunk_native(c, k, actions_host, obs_host, reward_host, done_host, status_host);
        if (rc) return rc;
    }
    if (c->wire == PZ_WIRE_COMPACT) wire_start_job(c, obs_host, reward_host, done_host, status_host);
    c->in_flight = true;
    return 0;
}

int pz_host_step_end(pz_host_ctx *c) {
    if (!c || !c->in_flight) return PZ_E_BADARG;
    DeviceGuard guard(c->device);
    c->in_flight = false;
    if (c->wire == PZ_WIRE_COMPACT) {
        const int rc = wire_finish_job(c);  // the workers have waited for every chunk's event
        const int rc2 = sync_all(c);
        return rc ? rc : rc2;
    }
    return sync_all(c);
}

int pz_host_step(pz_host_ctx *c, const void *actions_host, void *obs_host, void *reward_host,
                 uint8_t *done_host) {
    if (int rc = pz_host_step_begin(c, actions_host, obs_host, reward_host, done_host, nullptr)) return rc;
    return pz_host_step_end(c);
}

int pz_host_stats(pz_host_ctx *c, int64_t stats_host[PZ_NUM_STATS]) {
    if (!c || !stats_host || c->in_flight) return PZ_E_BADARG;
    DeviceGuard guard(c->device);
    int rc = sync_all(c);
    if (rc) return rc;
    PZ_CUDA(cudaMemcpy(stats_host, c->stats, PZ_NUM_STATS * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return 0;
}

int32_t *pz_host_state_dev(pz_host_ctx *c) { return c ? c->state : nullptr; }

void pz_host_destroy(pz_host_ctx *c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    for (cudaStream_t s : c->streams)
        if (s) cudaStreamSynchronize(s);
    wire_teardown(c);
    for (cudaStream_t s : c->streams)
        if (s) cudaStreamDestroy(s);
    cudaFree(c->state);
    cudaFree(c->actions);
    cudaFree(c->obs);
    cudaFree(c->reward);
    cudaFree(c->done);
    cudaFree(c->status);
    cudaFree(c->stats);
    delete c;
}

}  // extern "C"
