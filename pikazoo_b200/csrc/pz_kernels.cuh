// Internal launch interface shared by the C ABI (pz_kernels.cu) and the host-buffer path (pz_host.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pikazoo_b200.h"

namespace pz {

// 0, or PZ_E_ABI (handshake members of another revision) / PZ_E_BADCONFIG / PZ_E_BADARG (null)
int check_config(const pz_config *cfg);

// Step / reset the env range [begin, end) of a state buffer holding n envs (begin % 32 == 0).
// actions/obs/reward/done are the base pointers of the full [n]-sized arrays.
int launch_step(int32_t *state_dev, int64_t n, int64_t begin, int64_t end, const pz_config *cfg,
                const void *actions_dev, void *obs_dev, void *reward_dev, uint8_t *done_dev, int64_t *stats_dev,
                const pz_episode_io *episode, cudaStream_t stream);
int launch_reset(int32_t *state_dev, int64_t n, int64_t begin, int64_t end, const pz_config *cfg, void *obs_dev,
                 const pz_episode_io *episode, cudaStream_t stream);

// The observations of the state as it stands (obs_dev as pz_reset's).
int launch_observe(int32_t *state_dev, int64_t n, const pz_config *cfg, void *obs_dev, cudaStream_t stream);

}  // namespace pz
