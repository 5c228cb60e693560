// kernels.cuh — launch interface of the wavefront integrator (kernels.cu) used by abi.cu.
#pragma once
#ifdef VR_HOST_SHIM  // tests only (tests/c/host_shim.h)
#include "host_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "layout.h"

namespace vr {

// Wavefront state, structure-of-arrays over `capacity` path slots.
// Rays and hits of depth d are stored in QUEUE ORDER: entry i of the depth's compacted list sits at ray_o[d & 1][i],
// ray_d[d & 1][i], hit[i], and queue[d & 1][i] names the path slot it belongs to (depth 0 too: k_raygen may compact). k_trace streams
// its rays without an indirection and k_shade reads ray + hit coalesced; only the per-slot state (attenuation stack,
// finished radiance) is addressed through the slot.
struct Wavefront {
    float4* ray_o[2];  // origin.xyz, draws consumed so far (uint bits); ping-pong by depth parity
    float4* ray_d[2];  // direction.xyz (unit), unused
    float4* hit;     // t, GPU primitive index (int bits, -1 miss), u, v — queue order of the depth being traced
    float4* att;     // [max_bounces][capacity] attenuation of every level (rgb, unused), by slot (level-major: the
                     // writes of depth 0, whose queue is in slot order, coalesce; slot-major measured +40 % DRAM traffic)
    float4* radiance;  // [capacity] finished radiance of the slot's camera sample
    uint32_t* queue[2];  // compacted slot lists, ping-pong by depth parity
    uint32_t* counts;    // [max_bounces + 2] queue lengths
    uint32_t* cursors;   // [max_bounces + 2] next unclaimed queue entry (dynamic ray fetch); then miss_count
    unsigned long long* segments;  // scene.hit calls, whole render
    unsigned long long* culled;    // of those, camera rays k_raygen answered itself (they missed the scene's bounds)
    float4* miss;          // [capacity] misses of depth >= 1 waiting for k_miss: direction.xyz, slot | depth << 26
    uint32_t* miss_count;  // entries in `miss`, zeroed with the counts
    uint32_t capacity;     // < 2^26
};

// Division by a launch-invariant divisor without the ~25-instruction software divide: q = umulhi(x, m) >> s, exact for
// every x < 2^31 (m = floor(2^(31 + L) / d) + 1, L = ceil(log2 d), s = L - 1; d == 1 passes x through).
// tests/c/fastdiv_check.cpp sweeps it against the plain division.
struct FastDiv {
    uint32_t d, m, s;
};
inline FastDiv make_fast_div(uint32_t d) {
    FastDiv f{d, 0u, 0u};
    if (d <= 1u) return f;
    uint32_t L = 0;
    while ((1ull << L) < d) ++L;
    f.m = (uint32_t)((1ull << (31u + L)) / d) + 1u;
    f.s = L - 1u;
    return f;
}

// slot -> (pixel, global sample): implicit (slot % n_pixels, sample_base + slot / n_pixels) or explicit lists
struct PathSource {
    const uint32_t* pixel;   // null: implicit
    const uint32_t* sample;
    uint32_t n_pixels;
    uint32_t sample_base;
    uint32_t width, height;  // implicit mode enumerates pixels in 8x4 tiles
    FastDiv by_pixels, by_tiles_per_row;
};
inline PathSource make_path_source(const uint32_t* pixel, const uint32_t* sample, uint32_t width, uint32_t height,
                                   uint32_t sample_base) {
    PathSource src{};
    src.pixel = pixel;
    src.sample = sample;
    src.n_pixels = width * height;
    src.sample_base = sample_base;
    src.width = width;
    src.height = height;
    src.by_pixels = make_fast_div(src.n_pixels);
    src.by_tiles_per_row = make_fast_div((width & ~7u) >> 3);
    return src;
}

struct FrameParams {
    uint32_t width, height;
    int32_t pixel_mapping;
    uint32_t max_bounces;
    float firefly_clamp;
    int32_t render_mode;
    int32_t integrator;
    uint64_t seed;
};

struct LaunchDims {
    int sm_count;
    int trace_blocks_per_sm;
    int shade_blocks_per_sm;
    int shade_first_blocks_per_sm;
};
void query_launch_dims(LaunchDims* dims);

// cull: camera rays that miss the scene's bounds are finished by k_raygen itself and the depth-0 queue is compacted;
// without it entry i of the queue is slot i (gate kernels read hits by slot)
void launch_raygen(const DeviceScene& sc, const Wavefront& wf, const PathSource& src, const FrameParams& fp,
                   uint32_t n_paths, bool cull, const LaunchDims& ld, cudaStream_t stream);
void launch_trace(const DeviceScene& sc, const Wavefront& wf, uint32_t depth, uint32_t n_upper,
                  const LaunchDims& ld, cudaStream_t stream);
void launch_shade(const DeviceScene& sc, const Wavefront& wf, const PathSource& src, const FrameParams& fp,
                  uint32_t depth, uint32_t n_upper, const LaunchDims& ld, cudaStream_t stream);
// after the last depth: the environment lookups + unwinds that k_shade parked in wf.miss
void launch_miss(const DeviceScene& sc, const Wavefront& wf, const FrameParams& fp, uint32_t n_upper, const LaunchDims& ld,
                 cudaStream_t stream);
// partial[pixel] += sum over the batch's samples (in sample order); when `finish`, fold
// partial * (1/total_samples) into accum (alpha += alpha_inc: 1 per iterative_render call) and clear partial.
void launch_accumulate(const Wavefront& wf, float4* partial, float4* accum, uint32_t width, uint32_t height,
                       uint32_t samples_in_batch, int finish, float inv_total_samples, float alpha_inc, cudaStream_t stream);
void launch_resolve(const float4* accum, float4* out, uint32_t n_pixels, float scale, float exposure_mul,
                    float inv_gamma, int32_t tonemap, cudaStream_t stream);

// Peer-memory reduce: sum = own + peers[0] + peers[1] + ... (that order) per pixel, read with peer loads over
// NVLink. tonemap < 0: write the sum to `out` (may alias `own`); otherwise write the resolved pixel to `out`.
static const int MAX_PEERS = 15;
struct PeerList {
    const float4* ptr[MAX_PEERS];
    uint32_t n;
};
void launch_reduce_resolve_peers(const float4* own, PeerList peers, float4* out, uint32_t n_pixels, float scale,
                                 float exposure_mul, float inv_gamma, int32_t tonemap, cudaStream_t stream);

// RGB f32 (as uploaded) -> RGBA f32 texel records
void launch_expand_rgb(const float* rgb, float4* rgba, size_t n_texels, cudaStream_t stream);

// debug / gate kernels
void launch_trace_rays(const DeviceScene& sc, const float* origins, const float* dirs, uint64_t n,
                       uint32_t* surface, uint32_t* prim, float* t, cudaStream_t stream);
void launch_primary_ids(const DeviceScene& sc, const Wavefront& wf, uint32_t width, uint32_t height,
                        uint32_t* surface, uint32_t* prim, float* t, cudaStream_t stream);
void launch_rng_draws(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out, cudaStream_t stream);
void launch_unit_sphere(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out, cudaStream_t stream);
void launch_texture_sample(TextureRec tex, uint64_t n, const float* uv, float* rgb, cudaStream_t stream);
void launch_environment_sample(const DeviceScene& sc, uint64_t n, const float* dirs, float* rgb, cudaStream_t stream);

}  // namespace vr
