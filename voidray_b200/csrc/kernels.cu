// kernels.cu — sm_100a wavefront path-tracing kernels: ray generation, closest-hit BVH traversal
// (shared-memory short stack, Möller–Trumbore), shading with warp-aggregated queue compaction,
// accumulation and resolve/tonemap. Compiled with -fmad=false (see device_math.cuh).
//
// Reference path being replaced: voidray_renderer/src/render/iterative.rs:11-55 -> core/tracer.rs:9-56
// -> core/scene.rs:182-185 -> core/bvh.rs:132-160 -> core/mesh.rs:144-189 -> voidray_common/src/simple.rs,
// environments.rs -> shaders/post_process.glsl.
#include "device_math.cuh"
#include "kernels.cuh"

namespace vr {

static constexpr int TRACE_THREADS = 128;  // 64 x 16 and 256 x 4 measured level (profiles/r2_variants.md)
static constexpr int TRACE_MIN_BLOCKS = 8;  // 1024 threads x 64 registers = the SM's register file
static constexpr int SHADE_THREADS = 128;
#ifndef VR_SHADE_MIN_BLOCKS
#define VR_SHADE_MIN_BLOCKS 8
#endif
static constexpr int SHADE_MIN_BLOCKS = VR_SHADE_MIN_BLOCKS;  // <= 64 registers: k_shade is bound by memory latency x resident warps
}  // namespace vr
#include "traversal.cuh"
namespace vr {

// ------------------------------------------------------------------------------------------------
// Textures and environment (core/texture.rs:52-98, voidray_common/src/environments.rs:57-86)
// ------------------------------------------------------------------------------------------------
// (float)v / 255.0f, correctly rounded, without the division: one Newton step on v * (1 / 255) recovers the exact
// quotient for every v in 0..255 (checked exhaustively on the host: tests/c/unorm8_exact.c)
__device__ __forceinline__ float unorm8(uint32_t v) {
    const float x = (float)v, r = 1.0f / 255.0f;
    const float q = x * r;
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, x), r, q);
}
__device__ __forceinline__ f3 texel(const TextureRec& tex, uint32_t idx, uint32_t len) {
    if (idx >= len) idx -= len;  // `% len`: idx < 2*len always
    if (tex.pad == 1) {
        const uchar4 v = __ldg((const uchar4*)tex.texels + idx);
        return mk3(unorm8(v.x), unorm8(v.y), unorm8(v.z));
    }
    return xyz(ldg4((const float4*)tex.texels + idx));
}
__device__ __forceinline__ f3 bilinear_sample(const TextureRec& tex, float x, float y) {
    const uint32_t len = tex.width * tex.height;
    const uint32_t x0 = f32_as_index(x, tex.width - 1);
    const uint32_t y0 = f32_as_index(y, tex.height - 1);
    const float ax = x - (float)x0;
    const float ay = y - (float)y0;
    const f3 c00 = texel(tex, y0 * tex.width + x0, len);
    const f3 c01 = texel(tex, y0 * tex.width + x0 + 1, len);
    const f3 c10 = texel(tex, (y0 + 1) * tex.width + x0, len);
    const f3 c11 = texel(tex, (y0 + 1) * tex.width + x0 + 1, len);
    return lerp3(lerp3(c00, c01, ax), lerp3(c10, c11, ax), ay);
}
__device__ __forceinline__ f3 texture_sample(const TextureRec& tex, float u, float v) {
    if (u < 0.0f) u -= truncf(u) - 1.0f;
    if (v < 0.0f) v -= truncf(v) - 1.0f;
    const float x = fmodf(u, 1.0f) * (float)tex.width;
    const float y = (1.0f - fmodf(v, 1.0f)) * (float)tex.height;
    if (tex.sample_type == 0) {
        const uint32_t xi = f32_as_index(x, tex.width - 1);
        const uint32_t yi = f32_as_index(y, tex.height - 1);
        return texel(tex, yi * tex.width + xi, tex.width * tex.height);
    }
    return bilinear_sample(tex, x, y);
}
__device__ __forceinline__ f3 environment_sample(const DeviceScene& sc, f3 dir) {
    if (sc.env_kind == 1) return mk3(sc.env_color[0], sc.env_color[1], sc.env_color[2]);
    if (sc.env_kind != 2) return mk3(0.0f, 0.0f, 0.0f);
    const f3 d = normalize(dir);
    const float sx = acosf(-d.y);                   // util/math.rs:24-29
    const float sy = atan2f(-d.z, d.x) + VR_PI_F;
    const float u = sx / VR_PI_F;
    const float v = sy / (2.0f * VR_PI_F);
    const float x = v * (float)sc.env_tex.width;
    const float y = (float)(sc.env_tex.height - 1) - (u * (float)sc.env_tex.height);
    return bilinear_sample(sc.env_tex, x, y);
}

// ------------------------------------------------------------------------------------------------
// Ray generation: render/iterative.rs:25-42 + core/camera.rs:69-82
// ------------------------------------------------------------------------------------------------
// Within one sample the slots enumerate the frame in 8x4-pixel tiles, so the 32 primary rays of a warp
// cover a compact screen tile instead of a 32x1 strip (coherent node fetches at depth 0). Pixels right of
// the last full tile column / below the last full tile row follow in row order. Pure enumeration: the
// pixel index itself (and with it the camera mapping and the random stream) is unchanged.
__device__ __forceinline__ uint32_t fast_div(uint32_t x, const FastDiv& f) {  // x < 2^31, see kernels.cuh
    if (f.m == 0u) return f.d <= 1u ? x : x / f.d;  // d == 1, or no magic number prepared
    return __umulhi(x, f.m) >> f.s;
}
// (x, y) of slot-in-sample j; by_tpr divides by the tiles per row, (W & ~7) >> 3
__device__ __forceinline__ void tile_slot_to_xy(uint32_t j, uint32_t W, uint32_t H, const FastDiv& by_tpr, uint32_t& x, uint32_t& y) {
    const uint32_t W8 = W & ~7u, H4 = H & ~3u;
    const uint32_t n_tiled = W8 * H4;
    if (j < n_tiled) {
        const uint32_t tile = j >> 5, within = j & 31u;
        const uint32_t ty = fast_div(tile, by_tpr), tx = tile - ty * by_tpr.d;
        x = tx * 8 + (within & 7u);
        y = ty * 4 + (within >> 3);
        return;
    }
    uint32_t r = j - n_tiled;
    const uint32_t RW = W - W8;
    if (r < RW * H4) {
        y = r / RW;
        x = W8 + r % RW;
        return;
    }
    r -= RW * H4;
    y = H4 + r / W;
    x = r % W;
}
__device__ __forceinline__ uint32_t tile_slot_to_pixel(uint32_t j, uint32_t W, uint32_t H, const FastDiv& by_tpr) {
    uint32_t x, y;
    tile_slot_to_xy(j, W, H, by_tpr, x, y);
    return y * W + x;
}
// once per pixel (k_accumulate, gate kernels): the plain division
__device__ __forceinline__ uint32_t tile_slot_to_pixel(uint32_t j, uint32_t W, uint32_t H) {
    uint32_t x, y;
    tile_slot_to_xy(j, W, H, FastDiv{(W & ~7u) >> 3, 0u, 0u}, x, y);
    return y * W + x;
}
__device__ __forceinline__ uint32_t tile_pixel_to_slot(uint32_t pixel, uint32_t W, uint32_t H) {
    const uint32_t W8 = W & ~7u, H4 = H & ~3u;
    const uint32_t x = pixel % W, y = pixel / W;
    if (x < W8 && y < H4) return (((y >> 2) * (W8 >> 3) + (x >> 3)) << 5) + ((y & 3u) << 3) + (x & 7u);
    const uint32_t n_tiled = W8 * H4, RW = W - W8;
    if (y < H4) return n_tiled + y * RW + (x - W8);
    return n_tiled + RW * H4 + (y - H4) * W + x;
}

__device__ __forceinline__ void slot_source(const PathSource& src, uint32_t slot, uint32_t& pixel, uint32_t& sample) {
    if (src.pixel) {
        pixel = src.pixel[slot];
        sample = src.sample[slot];
    } else {
        const uint32_t s = fast_div(slot, src.by_pixels);
        pixel = tile_slot_to_pixel(slot - s * src.n_pixels, src.width, src.height, src.by_tiles_per_row);
        sample = src.sample_base + s;
    }
}

// ------------------------------------------------------------------------------------------------
// Closest-hit kernel: one ray per thread, queue of slots
// ------------------------------------------------------------------------------------------------
// Closest-hit kernel: persistent warps, one ray per lane, two mechanisms against SIMT divergence
// (ncu before: 6.8 of 32 threads active per instruction at depth >= 1, profiles/r1_trace_baseline.md):
//  1. dynamic ray replacement (Aila & Laine 2009): ray lengths inside a warp differ by an order of
//     magnitude once paths are incoherent, so a warp whose live lanes drop below REFILL_THRESHOLD pulls
//     fresh rays from the depth's queue (one atomicAdd per refill) instead of idling until its longest ray ends;
//  2. majority-vote stepping: a lane is either at an inner node or at a leaf. Each warp iteration executes
//     the state that holds more lanes (one vote each), the other lanes wait; waiting leaf lanes pile up
//     until they outvote the node lanes. At least half of the live lanes work in every iteration and, unlike
//     speculative traversal, no lane walks nodes that a pending leaf would have culled.
static constexpr int REFILL_THRESHOLD = 12;

// Vote parameters (measured sweep in profiles/README.md): the leaf step runs once the lanes waiting at a leaf exceed
// 1 / LEAF_VOTE_NUM of the lanes at inner nodes; one vote buys NODE_STEPS node steps or LEAF_STEPS triangle tests.
static constexpr int LEAF_VOTE_NUM = 2, NODE_STEPS = 4, LEAF_STEPS = 2;

__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_trace(DeviceScene sc, Wavefront wf, uint32_t depth, int refill_below) {
    __shared__ int s_stack[(SMEM_STACK + 1) * TRACE_THREADS];  // + the rays' scene-level visibility words
    const uint32_t n = wf.counts[depth];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(wf.segments, (unsigned long long)n);
    const float4* __restrict__ ray_o = (depth & 1u ? wf.ray_o[1] : wf.ray_o[0]);
    const float4* __restrict__ ray_d = (depth & 1u ? wf.ray_d[1] : wf.ray_d[0]);
    const float4* __restrict__ nodes = (const float4*)sc.nodes;
    const float4* __restrict__ tri_isect = (const float4*)sc.tri_isect;
    uint32_t* cursor = wf.cursors + depth;
    int* sstack = s_stack + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    Traversal tr;
    tr.cur = SENTINEL;
    bool have = false;
    bool exhausted = false;  // warp-uniform: the queue has no more rays
    uint32_t slot = 0;  // the ray's index in the depth's queue

    while (true) {
        if (!exhausted) {
            const unsigned need = __ballot_sync(0xFFFFFFFFu, !have);
            if (need) {
                const int leader = __ffs(need) - 1;
                const uint32_t cnt = (uint32_t)__popc(need);
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(cursor, cnt);
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if (!have) {
                    const uint32_t i = base + (uint32_t)__popc(need & lt_mask);
                    if (i < n) {
                        slot = i;  // queue order: the lanes that refill together read consecutive rays
                        const float4 ro = ld_once4(ray_o + i);
                        const float4 rd = ld_once4(ray_d + i);
                        trav_begin(tr, sc, xyz(ro), xyz(rd), sstack, TRACE_THREADS);
                        have = true;
                    }
                }
                if (base + cnt >= n) exhausted = true;
            }
        }
        if (!__any_sync(0xFFFFFFFFu, have)) break;
        while (true) {
            if (have && tr.cur == SENTINEL) {
                const HitResult h = trav_finish(tr, sc, sstack, TRACE_THREADS);
                wf.hit[slot] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
                have = false;
            }
            const bool at_node = have && is_inner(tr.cur);
            const bool at_leaf = have && tr.cur < 0;
            const unsigned m_node = __ballot_sync(0xFFFFFFFFu, at_node);
            const unsigned m_leaf = __ballot_sync(0xFFFFFFFFu, at_leaf);
            const int live = __popc(m_node | m_leaf);
            if (live == 0 || (!exhausted && live < refill_below)) break;
            const int n_node = __popc(m_node), n_leaf = __popc(m_leaf);
            // Vote: the leaf step runs once the lanes waiting at a leaf exceed 1/LEAF_VOTE_NUM of the lanes at
            // inner nodes; otherwise every lane that is (still) at an inner node takes NODE_STEPS node steps on
            // this one vote — the loop control costs ~25 full-width instructions, half a node step. Measured sweep
            // in profiles/README.md.
            if (n_node >= n_leaf * LEAF_VOTE_NUM) {
#pragma unroll
                for (int step = 0; step < NODE_STEPS; ++step) {
                    if (have && is_inner(tr.cur)) trav_node(tr, nodes, sstack, TRACE_THREADS);
                }
            } else {
#pragma unroll
                for (int step = 0; step < LEAF_STEPS; ++step) {
                    if (have && tr.cur < 0) trav_leaf_step(tr, sc, tri_isect, sstack, TRACE_THREADS);
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Shading: core/tracer.rs:19-56 unrolled over the wavefront. The reference recursion
//   L_d = 0 + min(att_d * L_{d+1}, clamp)      (tracer.rs:44-50, util/color.rs:30-36)
// is evaluated exactly: every level's attenuation is parked in wf.att and the product is unwound from
// the terminal value back to level 0 when the path ends, in the reference's own operation order.
// ------------------------------------------------------------------------------------------------
// ---- integrator 1 ("fast"; not in the reference) ----------------------------------------------------
// Density per solid angle of the reference's Lambertian direction normalize(n + UnitSphere) (simple.rs:116)
// for an arbitrary, possibly non-unit n: the points n + s lie on the unit sphere around n; a ray from the
// origin along w meets it at r = a c +- sqrt(a^2 c^2 - a^2 + 1) (a = |n|, c = cos(w, n)); projecting the uniform
// surface measure gives the sum over positive roots of r^2 / (4 pi |r - a c|). For |n| = 1 this is cos / pi.
__device__ __forceinline__ float lambert_reference_pdf(f3 w_unit, f3 n) {
    const float a = magnitude(n);
    if (!(a > 1.0e-6f)) return 1.0f / (4.0f * VR_PI_F);
    const float ac = dot(w_unit, n);
    const float disc = ac * ac - a * a + 1.0f;
    if (!(disc >= 0.0f)) return 0.0f;
    const float sq = fmaxf(sqrtf(disc), 1.0e-6f);
    const float r1 = ac + sq, r2 = ac - sq;
    float sum = 0.0f;
    if (r1 > 0.0f) sum += r1 * r1;
    if (r2 > 0.0f) sum += r2 * r2;
    return sum / (4.0f * VR_PI_F * sq);
}
// largest k in [0, n) with cdf[k] <= xi (cdf has n + 1 entries)
__device__ __forceinline__ uint32_t cdf_find(const float* __restrict__ cdf, uint32_t n, float xi) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (__ldg(cdf + mid) <= xi) lo = mid;
        else hi = mid;
    }
    return lo;
}
// HDRI luminance x sin(theta) table sampling; the texel <-> direction mapping is the lookup's own
// (environments.rs:80-86): x = phi / 2pi * W, y = (H - 1) - acos(-d.y) / pi * H.
__device__ __noinline__ f3 env_sample_direction(const DeviceScene& sc, Rng& rng) {
    const uint32_t W = sc.env_tex.width, H = sc.env_tex.height;
    const uint32_t j = cdf_find(sc.env_marginal, H, rng.v01());
    const uint32_t i = cdf_find(sc.env_cond + (size_t)j * (W + 1), W, rng.v01());
    const float x = (float)i + rng.v01();
    const float y = (float)j + rng.v01();
    float sx = ((float)(H - 1) - y) / (float)H * VR_PI_F;
    if (!(sx > 1.0e-6f)) sx = 1.0e-6f;
    const float ang = x / (float)W * (2.0f * VR_PI_F) - VR_PI_F;
    const float r = sinf(sx);
    return mk3(r * cosf(ang), -cosf(sx), -(r * sinf(ang)));
}
__device__ __noinline__ float env_pdf_direction(const DeviceScene& sc, f3 dir) {
    const uint32_t W = sc.env_tex.width, H = sc.env_tex.height;
    const f3 d = normalize(dir);
    const float sx = acosf(-d.y);
    const float sy = atan2f(-d.z, d.x) + VR_PI_F;
    const float x = sy / (2.0f * VR_PI_F) * (float)W;
    const float y = (float)(H - 1) - (sx / VR_PI_F * (float)H);
    if (!(y >= 0.0f)) return 0.0f;
    const uint32_t i = f32_as_index(x, W - 1);
    const uint32_t j = f32_as_index(y, H - 1);
    const float* row = sc.env_cond + (size_t)j * (W + 1);
    const float p_tex = (__ldg(sc.env_marginal + j + 1) - __ldg(sc.env_marginal + j)) * (__ldg(row + i + 1) - __ldg(row + i));
    return p_tex * ((float)W * (float)H) / (2.0f * VR_PI_F * VR_PI_F * fmaxf(sinf(sx), 1.0e-6f));
}

// ---- MicrofacetBSDF, voidray_common/src/microfacet.rs (same operation order as the reference) ----
__device__ __forceinline__ f3 lerp_v(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float powi2(float a) { return a * a; }
__device__ __forceinline__ float powi3(float a) { return a * (a * a); }  // __powisf2(a, 3): r = a; a *= a; r *= a
__device__ __forceinline__ bool sign_positive(float x) { return (__float_as_uint(x) >> 31) == 0u; }

// util/math.rs:50-60: Matrix3::new is column-major, so `local_to_world(n) * h` evaluates
// (ns . h, nss . h, n . h) — the transpose of a local-to-world basis. Reproduced as written.
__device__ __forceinline__ f3 local_to_world_mul(f3 normal, f3 h) {
    const bool is_normal = isfinite(normal.x) && fabsf(normal.x) >= 1.17549435e-38f;
    const f3 ns = is_normal ? normalize(mk3(normal.y, -normal.x, 0.0f)) : normalize(mk3(0.0f, -normal.z, normal.y));
    const f3 nss = cross(normal, ns);
    const f3 c0 = mk3(ns.x, nss.x, normal.x), c1 = mk3(ns.y, nss.y, normal.y), c2 = mk3(ns.z, nss.z, normal.z);
    return c0 * h.x + c1 * h.y + c2 * h.z;
}

__device__ __noinline__ f3 microfacet_bsdf(const MaterialRec& m, f3 n, f3 wo, f3 wi) {  // microfacet.rs:120-208
    const f3 color = mk3(m.color[0], m.color[1], m.color[2]);
    const f3 one = mk3(1.0f, 1.0f, 1.0f);
    const float n_dot_wi = dot(n, wi);
    const float n_dot_wo = dot(n, wo);
    const bool wi_outside = sign_positive(n_dot_wi);
    const bool wo_outside = sign_positive(n_dot_wo);
    if (!m.transparent && (!wi_outside || !wo_outside)) return mk3(0.0f, 0.0f, 0.0f);
    const float m2 = m.roughness * m.roughness;
    const float f0s = powi2((m.index - 1.0f) / (m.index + 1.0f));
    if (wi_outside == wo_outside) {
        const f3 h = normalize(wi + wo);
        const float wo_dot_h = dot(wo, h);
        const float n_dot_h = dot(n, h);
        const float nh2 = powi2(n_dot_h);
        const float dd = expf((nh2 - 1.0f) / (m2 * nh2)) / (m2 * VR_PI_F * nh2 * nh2);
        f3 f;
        if (!wi_outside && sqrtf(1.0f - wo_dot_h * wo_dot_h) * m.index > 1.0f) {
            f = one;
        } else {
            const f3 f0 = lerp_v(mk3(f0s, f0s, f0s), color, m.metallic);
            f = f0 + (one - f0) * powi5(1.0f - wo_dot_h);
        }
        float g = fminf(n_dot_wi * n_dot_h, n_dot_wo * n_dot_h);
        g = (2.0f * g) / wo_dot_h;
        g = fminf(g, 1.0f);
        const f3 specular = dd * f * g / (4.0f * n_dot_wo * n_dot_wi);
        if (m.transparent) return specular;
        const f3 diffuse = mul_elem(one - f, color) / VR_PI_F;
        return specular + diffuse;
    }
    const float eta_t = wo_outside ? m.index : 1.0f / m.index;
    const f3 h = normalize(wi * eta_t + wo);
    const float wi_dot_h = dot(wi, h);
    const float wo_dot_h = dot(wo, h);
    const float n_dot_h = dot(n, h);
    const float nh2 = powi2(n_dot_h);
    const float dd = expf((nh2 - 1.0f) / (m2 * nh2)) / (m2 * VR_PI_F * nh2 * nh2);
    const f3 f0 = lerp_v(mk3(f0s, f0s, f0s), color, m.metallic);
    const f3 f = f0 + (one - f0) * powi5(1.0f - fabsf(wi_dot_h));
    float g = fminf(fabsf(n_dot_wi * n_dot_h), fabsf(n_dot_wo * n_dot_h));
    g = (2.0f * g) / fabsf(wo_dot_h);
    g = fminf(g, 1.0f);
    const f3 btdf = fabsf(wi_dot_h * wo_dot_h / (n_dot_wi * n_dot_wo)) *
                    (dd * (one - f) * g / powi2(eta_t * wi_dot_h + wo_dot_h));
    return mul_elem(btdf, color);
}

__device__ __forceinline__ f3 microfacet_beckmann(f3 n, float m2, Rng& rng) {  // microfacet.rs:239-249
    const float theta = atanf(sqrtf(m2 * -logf(rng.gen_f32())));
    const float sin_t = sinf(theta), cos_t = cosf(theta);
    const f2 c = rng.unit_circle();
    return local_to_world_mul(n, mk3(c.x * sin_t, c.y * sin_t, cos_t));
}
__device__ __forceinline__ float microfacet_beckmann_pdf(f3 h, f3 n, float m2) {  // microfacet.rs:251-256
    const float cos_t = fabsf(dot(h, n));
    const float sin_t = sqrtf(1.0f - cos_t * cos_t);
    return (1.0f / (VR_PI_F * m2 * powi3(cos_t))) * expf(-powi2(sin_t / cos_t) / m2);
}

// microfacet.rs:222-313; false = None
__device__ __noinline__ bool microfacet_sample(const MaterialRec& m, f3 n, f3 wo, Rng& rng, f3& wi_out, float& pdf_out) {
    const float m2 = m.roughness * m.roughness;
    const float f0 = powi2((m.index - 1.0f) / (m.index + 1.0f));
    float f = (1.0f - m.metallic) * f0 + m.metallic * ((m.color[0] + m.color[1] + m.color[2]) / 3.0f);
    f = f * (1.0f - 0.2f) + 1.0f * 0.2f;
    const float eta_t = dot(wo, n) > 0.0f ? m.index : 1.0f / m.index;
    f3 wi;
    if (rng.gen_bool(f)) {
        const f3 h = microfacet_beckmann(n, m2, rng);
        wi = -reflect(wo, h);
    } else if (!m.transparent) {
        const f2 dsk = rng.unit_disc();
        const float z = sqrtf(1.0f - dsk.x * dsk.x - dsk.y * dsk.y);
        wi = local_to_world_mul(n, mk3(dsk.x, dsk.y, z));
    } else {
        const f3 h = microfacet_beckmann(n, m2, rng);
        const float cos_to = dot(h, wo);
        const f3 wo_perp = wo - h * cos_to;
        const f3 wi_perp = -wo_perp / eta_t;
        const float sin2_ti = magnitude2(wi_perp);
        if (sin2_ti > 1.0f) return false;
        const float cos_ti = sqrtf(1.0f - sin2_ti);
        const float sg = isnan(cos_to) ? cos_to : (sign_positive(cos_to) ? 1.0f : -1.0f);
        wi = -sg * cos_ti * h + wi_perp;
    }
    float p = 0.0f;
    {
        const f3 h = normalize(wi + wo);
        const float p_h = microfacet_beckmann_pdf(h, n, m2);
        p += f * p_h / (4.0f * fabsf(dot(h, wo)));
    }
    if (!m.transparent) {
        p += (1.0f - f) * fmaxf(dot(wi, n), 0.0f) / VR_PI_F;
    } else if (sign_positive(dot(wo, n)) != sign_positive(dot(wi, n))) {
        const f3 h = normalize(wi * eta_t + wo);
        const float p_h = microfacet_beckmann_pdf(h, n, m2);
        const float h_dot_wo = dot(h, wo);
        const float h_dot_wi = dot(h, wi);
        const float jacobian = fabsf(h_dot_wo) / powi2(eta_t * h_dot_wi + h_dot_wo);
        p += (1.0f - f) * p_h * jacobian;
    } else {
        p += 0.0f;
    }
    if (p == 0.0f) return false;
    wi_out = wi;
    pdf_out = p;
    return true;
}

__device__ __forceinline__ f3 clamp_color(f3 c, float mx) { return mk3(fminf(c.x, mx), fminf(c.y, mx), fminf(c.z, mx)); }

// L_level = value; fold levels level-1 .. 0
__device__ __forceinline__ f3 unwind(const Wavefront& wf, uint32_t slot, uint32_t level, f3 value, float clampv) {
    for (uint32_t d = level; d-- > 0;) {
        const f3 att = xyz(wf.att[(size_t)d * wf.capacity + slot]);
        value = mk3(0.0f, 0.0f, 0.0f) + clamp_color(mul_elem(att, value), clampv);
    }
    return value;
}

// FAST = integrator 1 compiled in, MICROFACET = the scene has a MicrofacetBSDF material; the common
// kernel (parity integrator, simple materials) carries neither code path.
// A miss: tracer.rs:30-33 returns the environment sample unclamped at this level; the path ends.
__device__ __forceinline__ void shade_miss(const DeviceScene& sc, const Wavefront& wf, float firefly_clamp, uint32_t depth,
                                           uint32_t slot, f3 d) {
    const f3 env = environment_sample(sc, d);
    const f3 L = unwind(wf, slot, depth, env, firefly_clamp);
    wf.radiance[slot] = make_float4(L.x, L.y, L.z, 0.0f);
}
// One path at one depth that hit something: its ray (ro, rd: queue-order records) and closest hit hr. Either the path
// ends here — its radiance is unwound and written, returns false — or it scatters: the level's attenuation is parked in
// wf.att, the next ray is returned in next_o / next_d, returns true.
template <bool FAST, bool MICROFACET>
__device__ __forceinline__ bool shade_hit(const DeviceScene& sc, const Wavefront& wf, const PathSource& src,
                                          const FrameParams& fp, uint32_t depth, uint32_t slot, float4 ro, float4 rd, float4 hr,
                                          float4& next_o, float4& next_d) {
    const float4* __restrict__ tri_shade = (const float4*)sc.tri_shade;
    bool alive = false;
    {
        {
            const f3 o = xyz(ro), d = xyz(rd);
            const int prim = __float_as_int(hr.y);
            {
                const float t = hr.x;
                const f3 point = o + d * t;  // Ray::at, util/ray.rs:19-21
                f3 outward;
                f2 uv;
                uint32_t material;
                if ((uint32_t)prim < sc.n_tris) {
                    const float u = hr.z, v = hr.w;
                    const float4 s0 = ldg4(tri_shade + 5 * prim), s1 = ldg4(tri_shade + 5 * prim + 1),
                                 s2 = ldg4(tri_shade + 5 * prim + 2), s3 = ldg4(tri_shade + 5 * prim + 3),
                                 s4 = ldg4(tri_shade + 5 * prim + 4);
                    const f3 n0 = xyz(s0), n1 = xyz(s1), n2 = xyz(s2), ng = xyz(s3);
                    const float w = 1.0f - u - v;
                    // mesh.rs:176-181
                    outward = u * n1 + v * n2 + w * n0;
                    uv.x = u * s2.w + v * s4.x + w * s0.w;
                    uv.y = u * s3.w + v * s4.y + w * s1.w;
                    if (angle_between(outward, ng) > 30.0f * VR_PI_F / 180.0f) outward = ng;
                    material = __float_as_uint(s4.z);
                } else {
                    const AnalyticRec a = sc.analytics[prim - sc.n_tris];
                    if (a.kind == 0) {
                        outward = (point - mk3(a.cx, a.cy, a.cz)) / a.radius;  // surfaces.rs:69-70
                        uv.x = 0.0f;
                        uv.y = 0.0f;
                    } else {
                        outward = mk3(0.0f, 1.0f, 0.0f);  // surfaces.rs:93-101
                        uv.x = point.x;
                        uv.y = point.z;
                    }
                    material = a.material;
                }
                // HitRecord::new, util/ray.rs:34-49
                const bool front_face = dot(d, outward) < 0.0f;
                const f3 normal = front_face ? outward : -outward;

                const MaterialRec m = sc.materials[material];
                f3 attenuation;
                bool scattered = false;
                f3 new_o = point, new_d = d;
                uint32_t pixel, sample;
                slot_source(src, slot, pixel, sample);
                Rng rng(fp.seed, pixel, sample, __float_as_uint(ro.w));

                if (fp.render_mode == 1) {
                    // tracer.rs:40: RenderMode::Normal
                    attenuation = 0.5f * normalize(normal) + mk3(1.0f, 1.0f, 1.0f) * 0.5f;
                } else if (FAST && m.kind == 0) {
                    // integrator 1: the integrand of Lambertian::scatter (albedo x the density of
                    // normalize(n + UnitSphere)), sampled by one-sample MIS with the HDRI luminance table
                    const f3 sn = m.normal_tex >= 0 ? texture_sample(sc.textures[m.normal_tex], uv.x, uv.y) : normal;
                    const f3 albedo = m.albedo_tex >= 0 ? texture_sample(sc.textures[m.albedo_tex], uv.x, uv.y)
                                                        : mk3(m.color[0], m.color[1], m.color[2]);
                    const bool has_env = sc.env_kind == 2;
                    f3 w;
                    if (has_env && rng.v01() >= 0.5f) {
                        w = normalize(env_sample_direction(sc, rng));
                    } else {
                        f3 dir = sn + rng.unit_sphere();
                        if (near_zero(dir)) dir = sn;
                        w = normalize(dir);
                    }
                    new_d = normalize(w);  // Ray::new
                    const float p_ref = lambert_reference_pdf(new_d, sn);
                    const float p_mix = has_env ? 0.5f * p_ref + 0.5f * env_pdf_direction(sc, new_d) : p_ref;
                    if (!(p_mix > 0.0f) || !(p_ref > 0.0f)) {
                        attenuation = mk3(0.0f, 0.0f, 0.0f);
                    } else {
                        attenuation = albedo * (p_ref / p_mix);
                        scattered = true;
                    }
                } else if (m.kind == 0) {
                    // Lambertian::scatter, simple.rs:103-132
                    const f3 sn = m.normal_tex >= 0 ? texture_sample(sc.textures[m.normal_tex], uv.x, uv.y) : normal;
                    f3 dir = sn + rng.unit_sphere();
                    if (near_zero(dir)) dir = sn;
                    new_d = normalize(dir);
                    scattered = true;
                    attenuation = m.albedo_tex >= 0 ? texture_sample(sc.textures[m.albedo_tex], uv.x, uv.y)
                                                    : mk3(m.color[0], m.color[1], m.color[2]);
                } else if (m.kind == 1) {
                    // Metal::scatter, simple.rs:141-160 (rejection loop bounded at 64 draws)
                    const f3 reflected = normalize(reflect(d, normal));
                    attenuation = mk3(m.color[0], m.color[1], m.color[2]);
                    for (int k = 0; k < 64; ++k) {
                        const f3 cand = normalize(reflected + m.param * rng.unit_sphere());
                        if (dot(cand, normal) > 0.0f) {
                            new_d = cand;
                            scattered = true;
                            break;
                        }
                    }
                    if (!scattered) attenuation = mk3(0.0f, 0.0f, 0.0f);  // gave up: the path ends BLACK
                } else if (m.kind == 2) {
                    // Dielectric::scatter, simple.rs:201-231
                    const float ratio = front_face ? 1.0f / m.param : m.param;
                    const f3 unit_direction = normalize(d);
                    const float cos_theta = fminf(dot(normal, -unit_direction), 1.0f);
                    const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
                    bool do_reflect = (ratio * sin_theta) > 1.0f;
                    if (!do_reflect) {
                        float r0 = (1.0f - ratio) / (1.0f + ratio);
                        r0 = r0 * r0;
                        const float refl = r0 + (1.0f - r0) * powi5(1.0f - cos_theta);
                        do_reflect = refl > rng.gen_range(0.0f, 1.0f);
                    }
                    const f3 dir = do_reflect ? reflect(unit_direction, normal) : refract(unit_direction, normal, ratio);
                    new_d = normalize(dir);
                    scattered = true;
                    attenuation = mk3(1.0f, 1.0f, 1.0f);
                } else if (m.kind == 3) {
                    // Emission::scatter, simple.rs:176-184 (colour * strength folded on the host)
                    attenuation = mk3(m.color[0], m.color[1], m.color[2]);
                } else if (MICROFACET && m.kind == 5) {
                    // MicrofacetBSDF through the blanket impl, core/traits.rs:23-40 + microfacet.rs:120-313
                    const f3 wo = normalize(d);  // the *incoming* direction, as in the reference
                    f3 wi;
                    float pdf;
                    attenuation = mk3(0.0f, 0.0f, 0.0f);
                    if (microfacet_sample(m, normal, wo, rng, wi, pdf)) {
                        const f3 f = microfacet_bsdf(m, normal, wo, wi);
                        attenuation = f * fabsf(dot(wi, normal)) * (1.0f / pdf);
                        new_d = normalize(wi);
                        scattered = true;
                    }
                } else {
                    // LambertianBSDF through the blanket impl, core/traits.rs:23-40 + simple.rs:64-81
                    const f3 wi = normalize(rng.unit_sphere());
                    const f3 f = mk3(m.color[0], m.color[1], m.color[2]) / VR_PI_F;
                    attenuation = f * fabsf(dot(wi, normal)) * (1.0f / 1.0f);
                    new_d = normalize(wi);
                    scattered = true;
                }

                // integrator 1: Russian roulette from the fourth segment on; survival probability = the largest
                // attenuation channel clamped to [0.05, 1], survivors are divided by it, the rest see BLACK
                bool killed = false;
                if (FAST && scattered && depth >= 3) {
                    const float q = fminf(fmaxf(fmaxf(attenuation.x, fmaxf(attenuation.y, attenuation.z)), 0.05f), 1.0f);
                    if (rng.v01() < q) attenuation = attenuation / q;
                    else killed = true;
                }

                if (!scattered) {
                    // tracer.rs:44-50 with no scattered ray: delta = attenuation
                    const f3 Lk = mk3(0.0f, 0.0f, 0.0f) + clamp_color(attenuation, fp.firefly_clamp);
                    const f3 L = unwind(wf, slot, depth, Lk, fp.firefly_clamp);
                    wf.radiance[slot] = make_float4(L.x, L.y, L.z, 0.0f);
                } else if (killed || depth + 1 >= fp.max_bounces) {
                    // the scattered ray would be traced at depth == max_bounces and return BLACK (tracer.rs:28)
                    const f3 Lk = mk3(0.0f, 0.0f, 0.0f) +
                                  clamp_color(mul_elem(attenuation, mk3(0.0f, 0.0f, 0.0f)), fp.firefly_clamp);
                    const f3 L = unwind(wf, slot, depth, Lk, fp.firefly_clamp);
                    wf.radiance[slot] = make_float4(L.x, L.y, L.z, 0.0f);
                } else {
                    wf.att[(size_t)depth * wf.capacity + slot] = make_float4(attenuation.x, attenuation.y, attenuation.z, 0.0f);
                    next_o = make_float4(new_o.x, new_o.y, new_o.z, __uint_as_float(rng.n));
                    next_d = make_float4(new_d.x, new_d.y, new_d.z, 0.0f);
                    alive = true;
                }
            }
        }
    }
    return alive;
}

// ------------------------------------------------------------------------------------------------
// Ray generation
// ------------------------------------------------------------------------------------------------
// One camera sample: render/iterative.rs:25-33 + core/camera.rs:58-82. (x, y) = the pixel centre on the film, jitter =
// 1 / the longer image side. Returns the ray and the draws consumed.
__device__ __forceinline__ void raygen_sample(const DeviceScene& sc, const FrameParams& fp, uint32_t pixel, uint32_t sample,
                                              float x, float y, float jitter, float4& ray_o, float4& ray_d) {
    Rng rng(fp.seed, pixel, sample, 0);
    const float dx = rng.gen_range(-jitter, jitter);
    const float dy = rng.gen_range(-jitter, jitter);
    const float cx = x + dx, cy = y + dy;

    const CameraRec& cam = sc.camera;
    const f3 right = mk3(cam.right[0], cam.right[1], cam.right[2]);
    const f3 up = mk3(cam.up[0], cam.up[1], cam.up[2]);
    f3 origin = mk3(cam.origin[0], cam.origin[1], cam.origin[2]);
    f3 new_dir = cam.d * mk3(cam.direction[0], cam.direction[1], cam.direction[2]) + cx * right + cy * up;
    if (cam.has_dof) {
        const f3 focal_point = origin + normalize(new_dir) * cam.focal_length;
        const f2 s = rng.unit_disc();
        origin = origin + (s.x * right + s.y * up) * cam.aperture;
        new_dir = focal_point - origin;
    }
    const f3 dir = normalize(normalize(new_dir));  // camera.rs:81 then Ray::new, util/ray.rs:15
    ray_o = make_float4(origin.x, origin.y, origin.z, __uint_as_float(rng.n));
    ray_d = make_float4(dir.x, dir.y, dir.z, 0.0f);
}
__device__ __forceinline__ void film_point(const FrameParams& fp, uint32_t px, uint32_t py, float dd, float& x, float& y) {
    x = ((float)(2u * px + 1u) - (float)fp.width) / dd;
    y = ((float)(2u * (fp.height - py) - 1u) - (float)fp.height) / dd;
}
// Does the ray miss everything for certain? The first step of k_trace's own traversal — the root record's two child
// boxes, same conservative slab test — so a ray culled here is a ray k_trace would have finished after one node step
// with no hit. Scenes with analytic surfaces (tested outside the tree) are never culled.
__device__ __forceinline__ bool misses_the_scene(const DeviceScene& sc, f3 o, f3 d) {
    if (sc.n_analytics != 0u) return false;
    if (sc.n_tris == 0u) return true;
    int stack[2];  // one push at most
    Traversal tr;
    trav_axis(o.x, d.x, sc.grid_min[0], sc.grid_extent[0], tr.ax, tr.bnx, tr.bfx, tr.selx);
    trav_axis(o.y, d.y, sc.grid_min[1], sc.grid_extent[1], tr.ay, tr.bny, tr.bfy, tr.sely);
    trav_axis(o.z, d.z, sc.grid_min[2], sc.grid_extent[2], tr.az, tr.bnz, tr.bfz, tr.selz);
    tr.best.t = INFINITY;
    tr.sp = 0;
    tr.cur = 0;
    trav_node(tr, (const float4*)sc.nodes, stack, 1);
    return tr.cur == SENTINEL;  // neither child box entered, nothing pushed
}

// Implicit source: a thread owns one pixel (slot-in-sample j, 8x4 tiles) and every `groups`-th sample of the batch, so
// the tile decode, the film point and their divisions are paid once per thread, not once per ray (the per-ray kernel
// was 12 % of config 1's step at 2.6x its own store bandwidth bound).
// cull != 0: a camera ray that misses the scene for certain (misses_the_scene) never enters the wavefront — its
// environment sample is its radiance (tracer.rs:30-33 at level 0), written here; the others are compacted into the
// depth-0 queue, one atomic per block and round. Such a ray skips a ray write, a k_trace refill, a hit record and a
// k_shade_first entry (+5.9 % / +4.8 % on configs 3 / 4, where about half of the frame is background); where nearly
// every camera ray enters the bounds the test and the block barriers only cost (-1.2 % on configs 2 / 5), so the host
// switches it off when a call culled less than 15 % of its camera rays (abi.cu). cull == 0 (and the gate kernels):
// every ray is queued, entry i = slot i.
__global__ void __launch_bounds__(256) k_raygen(DeviceScene sc, Wavefront wf, PathSource src, FrameParams fp,
                                                uint32_t n_paths, uint32_t groups, int cull) {
    __shared__ uint32_t s_warp_count[8], s_block_base;
    const uint32_t W = fp.width, H = fp.height;
    const float dd = (float)(W > H ? W : H);
    const float jitter = 1.0f / dd;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (src.pixel || !cull || fp.max_bounces == 0u) {  // every ray queued in slot order
        if (blockIdx.x == 0 && threadIdx.x == 0) wf.counts[0] = n_paths;
        const uint32_t samples = src.pixel ? 0u : n_paths / src.n_pixels;
        const uint32_t n_threads = src.pixel ? n_paths : src.n_pixels * groups;
        for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n_threads; id += gridDim.x * blockDim.x) {
            uint32_t pixel, px, py, g = 0, j = 0;
            if (src.pixel) {
                pixel = src.pixel[id];
                px = pixel % W;
                py = pixel / W;
            } else {
                g = fast_div(id, src.by_pixels);
                j = id - g * src.n_pixels;
                tile_slot_to_xy(j, W, H, src.by_tiles_per_row, px, py);
                pixel = py * W + px;
            }
            if (fp.pixel_mapping == 1) py = pixel / H;  // PixelMapping::Stretch quirk of the reference (iterative.rs:26)
            float x, y;
            film_point(fp, px, py, dd, x, y);
            for (uint32_t s = g; s < (src.pixel ? 1u : samples); s += groups) {
                const uint32_t slot = src.pixel ? id : s * src.n_pixels + j;
                float4 ro, rd;
                raygen_sample(sc, fp, pixel, src.pixel ? src.sample[id] : src.sample_base + s, x, y, jitter, ro, rd);
                wf.queue[0][slot] = slot;
                wf.ray_o[0][slot] = ro;
                wf.ray_d[0][slot] = rd;
                // every path writes its radiance exactly once, when it ends; with no bounce at all nothing does
                if (fp.max_bounces == 0u) wf.radiance[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        return;
    }
    const uint32_t n_pixels = src.n_pixels, samples = n_paths / n_pixels;
    const uint32_t n_threads = n_pixels * groups, rounds = (samples + groups - 1u) / groups;
    unsigned long long culled = 0;  // thread 0 of the block: camera rays answered here (they are segments all the same)
    // block-uniform loops: every thread of the block meets every __syncthreads
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_threads; base += gridDim.x * blockDim.x) {
        const uint32_t id = base + threadIdx.x;
        const bool owner = id < n_threads;
        uint32_t g = 0, j = 0, px = 0, py = 0, pixel = 0;
        float x = 0.0f, y = 0.0f;
        if (owner) {
            g = fast_div(id, src.by_pixels);
            j = id - g * n_pixels;
            tile_slot_to_xy(j, W, H, src.by_tiles_per_row, px, py);
            pixel = py * W + px;
            if (fp.pixel_mapping == 1) py = pixel / H;  // PixelMapping::Stretch quirk of the reference (iterative.rs:26)
            film_point(fp, px, py, dd, x, y);
        }
        for (uint32_t k = 0; k < rounds; ++k) {
            const uint32_t s = g + k * groups;
            const bool active = owner && s < samples;
            const uint32_t slot = s * n_pixels + j;
            float4 ro = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rd = ro;
            bool keep = false;
            if (active) {
                raygen_sample(sc, fp, pixel, src.sample_base + s, x, y, jitter, ro, rd);
                keep = !misses_the_scene(sc, xyz(ro), xyz(rd));
                if (!keep) shade_miss(sc, wf, fp.firefly_clamp, 0u, slot, xyz(rd));
            }
            // block-aggregated compaction into the depth-0 queue
            const unsigned keep_ballot = __ballot_sync(0xFFFFFFFFu, keep);
            const unsigned active_ballot = __ballot_sync(0xFFFFFFFFu, active);
            if (lane == 0u) s_warp_count[warp] = (uint32_t)__popc(keep_ballot) | ((uint32_t)__popc(active_ballot) << 16);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t kept = 0, seen = 0;
                for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
                    const uint32_t c = s_warp_count[w];
                    s_warp_count[w] = kept;  // the warp's offset inside the block's range
                    kept += c & 0xFFFFu;
                    seen += c >> 16;
                }
                culled += seen - kept;
                s_block_base = kept ? atomicAdd(&wf.counts[0], kept) : 0u;
            }
            __syncthreads();
            if (keep) {
                const uint32_t at = s_block_base + s_warp_count[warp] + (uint32_t)__popc(keep_ballot & ((1u << lane) - 1u));
                wf.queue[0][at] = slot;
                wf.ray_o[0][at] = ro;
                wf.ray_d[0][at] = rd;
            }
            __syncthreads();  // the counters are rewritten in the next round
        }
    }
    if (threadIdx.x == 0 && culled) {
        atomicAdd(wf.segments, culled);
        atomicAdd(wf.culled, culled);
    }
}

// Shading kernel. After the first bounce a warp's 32 queue entries are a mix of misses (environment lookup + unwind)
// and hits (triangle record, material, textures, scatter) — measured 27 miss / 5 hit lanes per instruction on config 1,
// 14 of 32 lanes overall (profiles/r2_trace_inst.md) — so the two are separated:
//  * a miss of depth >= 1 only appends {direction, slot, depth} to the batch's miss list; k_miss shades the whole list
//    after the last depth with every lane on the same code and, the list being in depth order, the same unwind length;
//  * hits are compacted per warp: a warp scans SHADE_SPAN consecutive 32-entry chunks at a time (all their loads in
//    flight together, one miss-list atomic per span), parks the indices of the hits in a ring in shared memory and runs
//    the material code on full groups of 32 (the rays and hit records it re-reads were touched a moment ago: L1 / L2).
// Depth 0 runs k_shade_first instead.
#ifndef VR_SHADE_SPAN
#define VR_SHADE_SPAN 4
#endif
static constexpr int SHADE_SPAN = VR_SHADE_SPAN;   // chunks a warp scans at a time
static constexpr uint32_t SHADE_RING = VR_SHADE_SPAN > 4 ? 512 : 256;  // >= 31 + 32 * SHADE_SPAN, power of two
static constexpr uint32_t MISS_SLOT_BITS = 26;  // miss record: slot | depth << 26 (capacity < 2^26, max_bounces <= 64)

template <bool FAST, bool MICROFACET>
__device__ __forceinline__ void shade_group(const DeviceScene& sc, const Wavefront& wf, const PathSource& src,
                                            const FrameParams& fp, uint32_t depth, const uint32_t* __restrict__ queue,
                                            uint32_t* __restrict__ queue_out, bool active, uint32_t i, uint32_t lane) {
    bool alive = false;
    uint32_t slot = 0;
    float4 next_o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), next_d = next_o;  // the scattered ray of a surviving path
    if (active) {
        slot = queue ? queue[i] : i;
        const float4 ro = (depth & 1u ? wf.ray_o[1] : wf.ray_o[0])[i];
        const float4 rd = (depth & 1u ? wf.ray_d[1] : wf.ray_d[0])[i];
        const float4 hr = wf.hit[i];
        if (__float_as_int(hr.y) < 0) shade_miss(sc, wf, fp.firefly_clamp, depth, slot, xyz(rd));  // depth 0 only
        else alive = shade_hit<FAST, MICROFACET>(sc, wf, src, fp, depth, slot, ro, rd, hr, next_o, next_d);
    }
    // warp-aggregated compaction: one atomic per warp
    __syncwarp();
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, alive);
    if (ballot) {
        uint32_t warp_base = 0;
        const int leader = __ffs(ballot) - 1;
        if ((int)lane == leader) warp_base = atomicAdd(&wf.counts[depth + 1], (uint32_t)__popc(ballot));
        warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, leader);
        if (alive) {
            // the next depth's ray goes to the path's place in the next queue (coalesced within the warp)
            const uint32_t j = warp_base + __popc(ballot & ((1u << lane) - 1u));
            queue_out[j] = slot;
            (depth & 1u ? wf.ray_o[0] : wf.ray_o[1])[j] = next_o;
            (depth & 1u ? wf.ray_d[0] : wf.ray_d[1])[j] = next_d;
        }
    }
}

// Depth 0: primary rays are coherent (hits and misses come in screen tiles) and a miss has nothing to unwind, so the
// first depth keeps the plain shape: one queue entry per thread, misses shaded in place (camera rays that miss the
// scene's bounds altogether may never get here: k_raygen).
template <bool FAST, bool MICROFACET>
__global__ void __launch_bounds__(SHADE_THREADS) k_shade_first(DeviceScene sc, Wavefront wf, PathSource src, FrameParams fp) {
    const uint32_t n = wf.counts[0];
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        shade_group<FAST, MICROFACET>(sc, wf, src, fp, 0u, wf.queue[0], wf.queue[1], i < n, i, lane);
    }
}

template <bool FAST, bool MICROFACET>
__global__ void __launch_bounds__(SHADE_THREADS, SHADE_MIN_BLOCKS) k_shade(DeviceScene sc, Wavefront wf, PathSource src,
                                                         FrameParams fp, uint32_t depth) {
    __shared__ uint32_t s_pending[SHADE_THREADS / 32][SHADE_RING];
    const uint32_t n = wf.counts[depth];
    const uint32_t* __restrict__ queue = depth == 0 ? nullptr : (depth & 1u ? wf.queue[1] : wf.queue[0]);
    uint32_t* __restrict__ queue_out = depth & 1u ? wf.queue[0] : wf.queue[1];
    const float4* __restrict__ ray_d = depth & 1u ? wf.ray_d[1] : wf.ray_d[0];
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t* pending = s_pending[threadIdx.x >> 5];
    uint32_t head = 0, count = 0;  // warp-uniform: the ring holds `count` indices from `head` on

    const uint32_t warps_per_block = blockDim.x >> 5;
    const uint32_t n_chunks = (n + 31u) >> 5;
    // the warp's spans: SHADE_SPAN consecutive chunks, then the span a whole grid further on
    uint32_t chunk = (blockIdx.x * warps_per_block + (threadIdx.x >> 5)) * SHADE_SPAN;
    const uint32_t chunk_stride = gridDim.x * warps_per_block * SHADE_SPAN;
    while (true) {
        while (count < 32u && chunk < n_chunks) {
            // every chunk's hit record of the span in flight at once
            int prim[SHADE_SPAN];
            unsigned miss_ballot[SHADE_SPAN];
            uint32_t n_miss = 0;
#pragma unroll
            for (int c = 0; c < SHADE_SPAN; ++c) {
                const uint32_t i = (chunk + c) * 32u + lane;
                prim[c] = i < n ? __float_as_int(wf.hit[i].y) : 0x7FFFFFFF;  // 0x7FFFFFFF: past the end
            }
#pragma unroll
            for (int c = 0; c < SHADE_SPAN; ++c) {
                const bool is_miss = prim[c] < 0;
                const bool is_hit = prim[c] != 0x7FFFFFFF && !is_miss;
                miss_ballot[c] = __ballot_sync(0xFFFFFFFFu, is_miss);
                n_miss += (uint32_t)__popc(miss_ballot[c]);
                const unsigned hit_ballot = __ballot_sync(0xFFFFFFFFu, is_hit);
                if (is_hit) pending[(head + count + (uint32_t)__popc(hit_ballot & lt_mask)) & (SHADE_RING - 1u)] = (chunk + c) * 32u + lane;
                count += (uint32_t)__popc(hit_ballot);
            }
            if (n_miss) {  // warp-uniform; one atomic for the span, the records in (chunk, lane) order
                uint32_t base = 0;
                if (lane == 0u) base = atomicAdd(wf.miss_count, n_miss);
                float4 rd[SHADE_SPAN];
                uint32_t sl[SHADE_SPAN];
#pragma unroll
                for (int c = 0; c < SHADE_SPAN; ++c) {
                    if ((miss_ballot[c] >> lane) & 1u) {
                        const uint32_t i = (chunk + c) * 32u + lane;
                        rd[c] = ray_d[i];
                        sl[c] = queue[i];
                    }
                }
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
                for (int c = 0; c < SHADE_SPAN; ++c) {
                    if ((miss_ballot[c] >> lane) & 1u)
                        wf.miss[base + (uint32_t)__popc(miss_ballot[c] & lt_mask)] =
                            make_float4(rd[c].x, rd[c].y, rd[c].z, __uint_as_float(sl[c] | (depth << MISS_SLOT_BITS)));
                    base += (uint32_t)__popc(miss_ballot[c]);
                }
            }
            chunk += chunk_stride;
        }
        if (count == 0u) break;  // the queue is exhausted and nothing is parked
        __syncwarp();
        const uint32_t take = count < 32u ? count : 32u;  // a partial group only at the very end
        const bool active = lane < take;
        const uint32_t idx = active ? pending[(head + lane) & (SHADE_RING - 1u)] : 0u;
        __syncwarp();
        shade_group<FAST, MICROFACET>(sc, wf, src, fp, depth, queue, queue_out, active, idx, lane);
        head = (head + take) & (SHADE_RING - 1u);
        count -= take;
    }
}

// The batch's misses of depth >= 1, in one launch after the last depth (see k_shade).
__global__ void __launch_bounds__(256) k_miss(DeviceScene sc, Wavefront wf, float firefly_clamp) {
    const uint32_t n = *wf.miss_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 rec = wf.miss[i];
        const uint32_t w = __float_as_uint(rec.w);
        shade_miss(sc, wf, firefly_clamp, w >> MISS_SLOT_BITS, w & ((1u << MISS_SLOT_BITS) - 1u), xyz(rec));
    }
}

// ------------------------------------------------------------------------------------------------
// Accumulation: render/iterative.rs:35-51
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_accumulate(Wavefront wf, float4* partial, float4* accum, uint32_t width,
                                                    uint32_t height, uint32_t samples_in_batch, int finish,
                                                    float inv_total, float alpha_inc) {
    const uint32_t n_pixels = width * height;
    // one thread per slot-in-sample j (coalesced radiance reads); it owns pixel tile_slot_to_pixel(j)
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_pixels; j += gridDim.x * blockDim.x) {
        const uint32_t px = tile_slot_to_pixel(j, width, height);
        float4 p = partial[px];
        for (uint32_t s = 0; s < samples_in_batch; ++s) {
            const float4 L = wf.radiance[(size_t)s * n_pixels + j];
            p.x += L.x;
            p.y += L.y;
            p.z += L.z;
        }
        if (finish) {
            float4 a = accum[px];
            a.x += p.x * inv_total;
            a.y += p.y * inv_total;
            a.z += p.z * inv_total;
            a.w += alpha_inc;  // Color::a() == 1.0 per call, util/color.rs:54 / iterative.rs:51 (0 on a group's other shards)
            accum[px] = a;
            p = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        partial[px] = p;
    }
}

// ------------------------------------------------------------------------------------------------
// Resolve: shaders/post_process.glsl:23-49 + shaders/tonemapping.glsl:2-40 (GLSL mat3 is column-major)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 mat3_mul(const float* m, f3 v) {
    return mk3(m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
               m[2] * v.x + m[5] * v.y + m[8] * v.z);
}
__device__ __forceinline__ float aces_fit(float c) {
    return (c * (c + 0.0245786f) - 0.000090537f) / (c * (0.983729f * c + 0.432951f) + 0.238081f);
}
__device__ __forceinline__ float filmic_fit(float c) { return (c * (6.2f * c + 0.5f)) / (c * (6.2f * c + 1.7f) + 0.06f); }
__device__ __forceinline__ float uncharted_fit(float c) {
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((c * (A * c + C * B) + D * E) / (c * (A * c + B) + D * F)) - E / F;
}

__device__ __forceinline__ float4 resolve_pixel(float4 a, float scale, float exposure_mul, float inv_gamma, int32_t tonemap) {
    const float ACES_IN[9] = {0.59719f, 0.076f, 0.0284f, 0.35458f, 0.90834f, 0.13383f, 0.04823f, 0.01566f, 0.83777f};
    const float ACES_OUT[9] = {1.60475f, -0.10208f, -0.00327f, -0.53108f, 1.10813f, -0.07276f, -0.07367f, -0.00605f, 1.07602f};
    f3 c = mk3(a.x * scale, a.y * scale, a.z * scale) * exposure_mul;
    if (tonemap == 1) {
        c = mat3_mul(ACES_IN, c);
        c = mk3(aces_fit(c.x), aces_fit(c.y), aces_fit(c.z));
        c = mat3_mul(ACES_OUT, c);
    } else if (tonemap == 2) {
        const float white = 2.0f;
        const float luma = (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f;
        const float tm = luma * (1.0f + luma / (white * white)) / (1.0f + luma);
        c = c * (tm / luma);
    } else if (tonemap == 3) {
        c = mk3(fmaxf(0.0f, c.x - 0.004f), fmaxf(0.0f, c.y - 0.004f), fmaxf(0.0f, c.z - 0.004f));
        c = mk3(filmic_fit(c.x), filmic_fit(c.y), filmic_fit(c.z));
    } else if (tonemap == 4) {
        c = c * 2.0f;
        c = mk3(uncharted_fit(c.x), uncharted_fit(c.y), uncharted_fit(c.z));
        c = c / uncharted_fit(11.2f);
    }
    return make_float4(powf(c.x, inv_gamma), powf(c.y, inv_gamma), powf(c.z, inv_gamma), 1.0f);
}

__global__ void __launch_bounds__(256) k_resolve(const float4* __restrict__ accum, float4* __restrict__ out,
                                                 uint32_t n_pixels, float scale, float exposure_mul, float inv_gamma,
                                                 int32_t tonemap) {
    for (uint32_t px = blockIdx.x * blockDim.x + threadIdx.x; px < n_pixels; px += gridDim.x * blockDim.x)
        out[px] = resolve_pixel(accum[px], scale, exposure_mul, inv_gamma, tonemap);
}

// Fused reduce + resolve over peer memory: every peer pointer is another GPU's accumulation buffer mapped into this
// address space (CUDA IPC / peer access), so the loads below travel over NVLink and the sum never exists as a
// separate buffer. All peer loads of a pixel are issued before the first add (independent 16-byte loads in flight).
__global__ void __launch_bounds__(256) k_reduce_resolve_peers(const float4* own, PeerList peers, float4* out, uint32_t n_pixels,
                                                              float scale, float exposure_mul, float inv_gamma, int32_t tonemap) {
    for (uint32_t px = blockIdx.x * blockDim.x + threadIdx.x; px < n_pixels; px += gridDim.x * blockDim.x) {
        float4 v[MAX_PEERS];
#pragma unroll
        for (int k = 0; k < MAX_PEERS; ++k)
            if (k < (int)peers.n) v[k] = peers.ptr[k][px];
        float4 a = own[px];
#pragma unroll
        for (int k = 0; k < MAX_PEERS; ++k)
            if (k < (int)peers.n) {
                a.x += v[k].x;
                a.y += v[k].y;
                a.z += v[k].z;
                a.w += v[k].w;
            }
        out[px] = tonemap < 0 ? a : resolve_pixel(a, scale, exposure_mul, inv_gamma, tonemap);
    }
}

// RGB f32 texels as uploaded -> the 16-byte RGBA records the lookups fetch (w = 0)
__global__ void __launch_bounds__(256) k_expand_rgb(const float* __restrict__ rgb, float4* __restrict__ rgba, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        rgba[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 0.0f);
}

// ------------------------------------------------------------------------------------------------
// Gate / debug kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void export_hit(const DeviceScene& sc, const HitResult& h, uint32_t* surface, uint32_t* prim,
                                           float* t, uint64_t i) {
    if (h.prim < 0) {
        surface[i] = 0xFFFFFFFFu;
        prim[i] = 0xFFFFFFFFu;
        t[i] = INFINITY;
    } else if ((uint32_t)h.prim < sc.n_tris) {
        surface[i] = sc.tri_surface[h.prim];
        prim[i] = sc.tri_prim[h.prim];
        t[i] = h.t;
    } else {
        surface[i] = sc.analytics[h.prim - sc.n_tris].surface;
        prim[i] = 0xFFFFFFFFu;
        t[i] = h.t;
    }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_rays(DeviceScene sc, const float* origins, const float* dirs,
                                                              uint64_t n, uint32_t* surface, uint32_t* prim, float* t) {
    __shared__ int s_stack[(SMEM_STACK + 1) * TRACE_THREADS];  // + the rays' scene-level visibility words
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const f3 o = mk3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        const f3 d = normalize(mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));  // Ray::new
        const HitResult h = closest_hit(sc, o, d, s_stack + threadIdx.x, TRACE_THREADS);
        export_hit(sc, h, surface, prim, t, i);
    }
}

__global__ void __launch_bounds__(256) k_primary_ids(DeviceScene sc, Wavefront wf, uint32_t n, uint32_t width,
                                                     uint32_t height, uint32_t* surface, uint32_t* prim, float* t) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 hr = wf.hit[i];
        HitResult h;
        h.t = hr.x;
        h.prim = __float_as_int(hr.y);
        h.u = hr.z;
        h.v = hr.w;
        export_hit(sc, h, surface, prim, t, tile_slot_to_pixel(i, width, height));
    }
}

__global__ void k_rng_draws(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Rng rng(seed, pixel, sample, 0);
        for (uint32_t i = 0; i < n; ++i) out[i] = rng.next_u32();
    }
}
__global__ void k_unit_sphere(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Rng rng(seed, pixel, sample, 0);
        for (uint32_t i = 0; i < n; ++i) {
            const f3 v = rng.unit_sphere();
            out[3 * i] = v.x;
            out[3 * i + 1] = v.y;
            out[3 * i + 2] = v.z;
        }
    }
}
__global__ void k_texture_sample(TextureRec tex, uint64_t n, const float* uv, float* rgb) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const f3 c = texture_sample(tex, uv[2 * i], uv[2 * i + 1]);
        rgb[3 * i] = c.x;
        rgb[3 * i + 1] = c.y;
        rgb[3 * i + 2] = c.z;
    }
}
__global__ void k_environment_sample(DeviceScene sc, uint64_t n, const float* dirs, float* rgb) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const f3 c = environment_sample(sc, mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        rgb[3 * i] = c.x;
        rgb[3 * i + 1] = c.y;
        rgb[3 * i + 2] = c.z;
    }
}

#ifndef VR_HOST_SHIM  // (tests/c/ktrace_host.cpp compiles the kernels above for the CPU and launches them itself)
// ------------------------------------------------------------------------------------------------
// Launch wrappers. Grids are persistent-style: a multiple of the SM count x resident blocks, with
// grid-stride loops; the live queue length is read on the device, so no host round trip per depth.
// ------------------------------------------------------------------------------------------------
void query_launch_dims(LaunchDims* dims) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&dims->sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dims->trace_blocks_per_sm, k_trace, TRACE_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dims->shade_blocks_per_sm, k_shade<false, false>, SHADE_THREADS, 0);
    if (dims->trace_blocks_per_sm < 1) dims->trace_blocks_per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dims->shade_first_blocks_per_sm, k_shade_first<false, false>, SHADE_THREADS, 0);
    if (dims->shade_first_blocks_per_sm < 1) dims->shade_first_blocks_per_sm = 1;
}

static inline uint32_t grid_for(uint64_t n, int threads, int sm_count, int blocks_per_sm) {
    const uint64_t need = (n + threads - 1) / threads;
    const uint64_t cap = (uint64_t)sm_count * blocks_per_sm;
    const uint64_t g = need < cap ? need : cap;
    return (uint32_t)(g < 1 ? 1 : g);
}

void launch_raygen(const DeviceScene& sc, const Wavefront& wf, const PathSource& src, const FrameParams& fp,
                   uint32_t n_paths, bool cull, const LaunchDims& ld, cudaStream_t stream) {
    // enough threads for ~4 blocks of 256 per SM, at most one per ray
    uint32_t groups = 1;
    if (!src.pixel) {
        const uint32_t samples = n_paths / src.n_pixels;
        const uint64_t want = (uint64_t)ld.sm_count * 8 * 256;
        groups = (uint32_t)((want + src.n_pixels - 1) / src.n_pixels);
        groups = groups < 1 ? 1 : (groups > samples ? samples : groups);
        if (groups < 1) groups = 1;
    }
    const uint64_t threads = src.pixel ? n_paths : (uint64_t)src.n_pixels * groups;
    k_raygen<<<grid_for(threads, 256, ld.sm_count, 8), 256, 0, stream>>>(sc, wf, src, fp, n_paths, groups, cull ? 1 : 0);
}
void launch_trace(const DeviceScene& sc, const Wavefront& wf, uint32_t depth, uint32_t n_upper, const LaunchDims& ld,
                  cudaStream_t stream) {
    k_trace<<<grid_for(n_upper, TRACE_THREADS, ld.sm_count, ld.trace_blocks_per_sm), TRACE_THREADS, 0, stream>>>(
        sc, wf, depth, REFILL_THRESHOLD);
}
void launch_shade(const DeviceScene& sc, const Wavefront& wf, const PathSource& src, const FrameParams& fp,
                  uint32_t depth, uint32_t n_upper, const LaunchDims& ld, cudaStream_t stream) {
    const bool fast = fp.integrator == 1, mf = sc.has_microfacet != 0;
    if (depth == 0) {
        const uint32_t grid = grid_for(n_upper, SHADE_THREADS, ld.sm_count, ld.shade_first_blocks_per_sm);
        if (fast && mf) k_shade_first<true, true><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp);
        else if (fast) k_shade_first<true, false><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp);
        else if (mf) k_shade_first<false, true><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp);
        else k_shade_first<false, false><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp);
        return;
    }
    const uint32_t grid = grid_for(n_upper, SHADE_THREADS, ld.sm_count, ld.shade_blocks_per_sm);
    if (fast && mf) k_shade<true, true><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp, depth);
    else if (fast) k_shade<true, false><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp, depth);
    else if (mf) k_shade<false, true><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp, depth);
    else k_shade<false, false><<<grid, SHADE_THREADS, 0, stream>>>(sc, wf, src, fp, depth);
}
void launch_miss(const DeviceScene& sc, const Wavefront& wf, const FrameParams& fp, uint32_t n_upper, const LaunchDims& ld,
                 cudaStream_t stream) {
    k_miss<<<grid_for(n_upper, 256, ld.sm_count, 8), 256, 0, stream>>>(sc, wf, fp.firefly_clamp);
}
void launch_accumulate(const Wavefront& wf, float4* partial, float4* accum, uint32_t width, uint32_t height,
                       uint32_t samples_in_batch, int finish, float inv_total_samples, float alpha_inc, cudaStream_t stream) {
    const uint32_t grid = (width * height + 255) / 256;
    k_accumulate<<<grid, 256, 0, stream>>>(wf, partial, accum, width, height, samples_in_batch, finish, inv_total_samples,
                                           alpha_inc);
}
void launch_resolve(const float4* accum, float4* out, uint32_t n_pixels, float scale, float exposure_mul, float inv_gamma,
                    int32_t tonemap, cudaStream_t stream) {
    const uint32_t grid = (n_pixels + 255) / 256;
    k_resolve<<<grid, 256, 0, stream>>>(accum, out, n_pixels, scale, exposure_mul, inv_gamma, tonemap);
}
void launch_expand_rgb(const float* rgb, float4* rgba, size_t n_texels, cudaStream_t stream) {
    if (n_texels == 0) return;
    const size_t grid = (n_texels + 255) / 256;
    k_expand_rgb<<<(unsigned)(grid > 148 * 16 ? 148 * 16 : grid), 256, 0, stream>>>(rgb, rgba, n_texels);
}
void launch_reduce_resolve_peers(const float4* own, PeerList peers, float4* out, uint32_t n_pixels, float scale,
                                 float exposure_mul, float inv_gamma, int32_t tonemap, cudaStream_t stream) {
    const uint32_t grid = (n_pixels + 255) / 256;
    k_reduce_resolve_peers<<<grid > 148 * 8 ? 148 * 8 : grid, 256, 0, stream>>>(own, peers, out, n_pixels, scale,
                                                                              exposure_mul, inv_gamma, tonemap);
}
void launch_trace_rays(const DeviceScene& sc, const float* origins, const float* dirs, uint64_t n, uint32_t* surface,
                       uint32_t* prim, float* t, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)((n + TRACE_THREADS - 1) / TRACE_THREADS);
    k_trace_rays<<<grid < 1 ? 1 : (grid > 148 * 8 ? 148 * 8 : grid), TRACE_THREADS, 0, stream>>>(sc, origins, dirs, n, surface, prim, t);
}
void launch_primary_ids(const DeviceScene& sc, const Wavefront& wf, uint32_t width, uint32_t height, uint32_t* surface,
                        uint32_t* prim, float* t, cudaStream_t stream) {
    const uint32_t n = width * height;
    k_primary_ids<<<(n + 255) / 256, 256, 0, stream>>>(sc, wf, n, width, height, surface, prim, t);
}
void launch_rng_draws(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out, cudaStream_t stream) {
    k_rng_draws<<<1, 32, 0, stream>>>(seed, pixel, sample, n, out);
}
void launch_unit_sphere(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out, cudaStream_t stream) {
    k_unit_sphere<<<1, 32, 0, stream>>>(seed, pixel, sample, n, out);
}
void launch_texture_sample(TextureRec tex, uint64_t n, const float* uv, float* rgb, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)((n + 255) / 256);
    k_texture_sample<<<grid < 1 ? 1 : (grid > 4096 ? 4096 : grid), 256, 0, stream>>>(tex, n, uv, rgb);
}
void launch_environment_sample(const DeviceScene& sc, uint64_t n, const float* dirs, float* rgb, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)((n + 255) / 256);
    k_environment_sample<<<grid < 1 ? 1 : (grid > 4096 ? 4096 : grid), 256, 0, stream>>>(sc, n, dirs, rgb);
}

#endif  // VR_HOST_SHIM

}  // namespace vr
