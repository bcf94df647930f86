// pz_render: rasterises display lists into rgb_array frames — the pixel work of raw_env.render() / draw()
// (pikazoo/env/pikazoo_env.py:250-384) for a batch of selected envs. The host side (pikazoo_b200/render.py) turns
// simulation states into display lists in the reference's draw order (pinned against the reference's own draw());
// here one thread owns one output pixel and walks its frame's list from the front-most item backwards, stopping at
// the first opaque texel. Every sprite of the reference has binary alpha, so compositing is a select and the result
// does not depend on a blend formula. The static background (446 blits, identical every frame) is composited once
// on the host and arrives as an image. Not a hot path: 131,328 pixels x <= 64 bounding-box tests per frame.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pikazoo_b200.h"

namespace pzr {

constexpr int kW = 432, kH = 304;  // GROUND_WIDTH, GROUND_HEIGHT (pikazoo_env.py:24)

__global__ void __launch_bounds__(256) pz_render_kernel(const uchar4 *__restrict__ atlas, const int4 *__restrict__ sprites,
                                                        int n_sprites, const unsigned char *__restrict__ background,
                                                        const int4 *__restrict__ items, int max_items,
                                                        unsigned char *__restrict__ out) {
    extern __shared__ int4 s_items[];  // this frame's list: (sprite variant, x, y, -) ; variant < 0 = unused slot
    const int frame = blockIdx.y;
    for (int m = threadIdx.x; m < max_items; m += blockDim.x) s_items[m] = items[(size_t)frame * max_items + m];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kW * kH) return;
    const int x = p % kW, y = p / kW;
    unsigned char r = background[3 * p], g = background[3 * p + 1], b = background[3 * p + 2];
    for (int m = max_items - 1; m >= 0; m--) {
        const int4 it = s_items[m];
        if (it.x < 0 || it.x >= n_sprites) continue;
        const int4 sp = sprites[it.x];  // (offset in texels, w, h, -)
        const int dx = x - it.y, dy = y - it.z;
        if ((unsigned)dx >= (unsigned)sp.y || (unsigned)dy >= (unsigned)sp.z) continue;
        const uchar4 t = atlas[sp.x + dy * sp.y + dx];
        if (t.w != 0) {
            r = t.x, g = t.y, b = t.z;
            break;
        }
    }
    unsigned char *o = out + ((size_t)frame * kW * kH + p) * 3;
    o[0] = r, o[1] = g, o[2] = b;
}

}  // namespace pzr

extern "C" int pz_render(const uint8_t *atlas_dev, const int32_t *sprites_dev, int32_t n_sprites,
                         const uint8_t *background_dev, const int32_t *items_dev, int32_t n_frames, int32_t max_items,
                         uint8_t *out_dev, void *stream) {
    if (!atlas_dev || !sprites_dev || !background_dev || !items_dev || !out_dev || n_sprites < 1 || n_frames < 0 ||
        max_items < 1 || max_items > 1024)
        return PZ_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(atlas_dev) & 3u) || (reinterpret_cast<uintptr_t>(sprites_dev) & 15u) ||
        (reinterpret_cast<uintptr_t>(items_dev) & 15u))
        return PZ_E_ALIGN;
    if (n_frames == 0) return 0;
    const dim3 grid((pzr::kW * pzr::kH + 255) / 256, (unsigned)n_frames);
    pzr::pz_render_kernel<<<grid, 256, (size_t)max_items * sizeof(int4), (cudaStream_t)stream>>>(
        reinterpret_cast<const uchar4 *>(atlas_dev), reinterpret_cast<const int4 *>(sprites_dev), n_sprites, background_dev,
        reinterpret_cast<const int4 *>(items_dev), max_items, out_dev);
    cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? 0 : (int)err;
}
