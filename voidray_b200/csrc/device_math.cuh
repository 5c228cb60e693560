// device_math.cuh — f32 vector arithmetic in the operation order of cgmath 0.18 (the reference's
// maths crate), the counter-based generator, and the rand / rand_distr distribution algorithms.
//
// This translation unit is compiled with -fmad=false: every `a*b+c` below is two IEEE roundings, in
// source order, exactly like rustc emits for the reference. That is what makes hit distances,
// barycentrics, normals and scatter directions bit-identical to a CPU evaluation of the same
// expressions; FMA is used only where written explicitly (__fmaf_rn).
#pragma once
#ifdef VR_HOST_SHIM  // tests only: this source compiled for the CPU (tests/c/host_shim.h, tests/c/trav_host.cpp)
#include "host_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace vr {

struct f3 {
    float x, y, z;
};
struct f2 {
    float x, y;
};

__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return f3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return f3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ f3 mul_elem(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float magnitude2(f3 a) { return dot(a, a); }
__device__ __forceinline__ float magnitude(f3 a) { return sqrtf(magnitude2(a)); }
__device__ __forceinline__ f3 normalize(f3 a) { return a * (1.0f / magnitude(a)); }
// cgmath Vector3::angle
__device__ __forceinline__ float angle_between(f3 a, f3 b) { return atan2f(magnitude(cross(a, b)), dot(a, b)); }

__device__ __forceinline__ f3 xyz(float4 q) { return f3{q.x, q.y, q.z}; }
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only load (PTX ISA 8.8, sm_100+: SASS LDG.E.ENL2.256.CONSTANT). One instruction per
// 32 bytes halves the L1 tag traffic of the scattered node / triangle fetches; `p` must be 32-byte aligned.
struct __align__(32) float8 {
    float4 lo, hi;
};
__device__ __forceinline__ float8 ldg8(const float4* p) {
    float8 r;
#ifdef VR_HOST_SHIM
    r.lo = p[0];
    r.hi = p[1];
#else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
#endif
    return r;
}

// The same 256-bit load with an L1 policy. The SM's L1 (what the eight blocks' stacks leave of 256 KB) cannot hold a
// scene's nodes and triangles, so what it keeps matters: a node is re-used by every ray of the SM (evict_last), a
// triangle record by few (no_allocate: it does not push nodes out). Measured +1.2 ... +1.9 % on every config, with
// the rays' own loads bypassing the L1 as well +1.0 ... +2.3 % (profiles/r2_variants.md, calls 16-19).
__device__ __forceinline__ float8 ldg8_keep(const float4* p) {
#ifndef VR_HOST_SHIM
    float8 r;
    asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
    return r;
#else
    return ldg8(p);
#endif
}
__device__ __forceinline__ float8 ldg8_once(const float4* p) {
#ifndef VR_HOST_SHIM
    float8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                   "=f"(r.hi.w)
                 : "l"(p));
    return r;
#else
    return ldg8(p);
#endif
}

// A ray is read once by k_trace: its two quads bypass the L1
__device__ __forceinline__ float4 ld_once4(const float4* p) {
#ifndef VR_HOST_SHIM
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return *p;
#endif
}

#define VR_PI_F 3.14159265358979323846f

// compiler-rt __powisf2 (what f32::powi lowers to), specialised for exponent 5: a * (a^2)^2
__device__ __forceinline__ float powi5(float a) {
    const float a2 = a * a;
    const float a4 = a2 * a2;
    return a * a4;
}

// Rust `x as usize` then `.min(limit)`: saturating, NaN -> 0
__device__ __forceinline__ uint32_t f32_as_index(float x, uint32_t limit) {
    const uint32_t i = __float2uint_rz(x);  // saturates, NaN -> 0
    return i < limit ? i : limit;
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One camera sample's draw stream: draw i is word (i & 3) of Philox(counter = (pixel, sample, i >> 2, 0)).
// Only `n` (draws consumed) has to survive between wavefront stages.
struct Rng {
    uint32_t k0, k1, pixel, sample, n;
    uint32_t b0, b1, b2, b3, block;
    __device__ __forceinline__ Rng(uint64_t seed, uint32_t pixel_, uint32_t sample_, uint32_t n_)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), pixel(pixel_), sample(sample_), n(n_), b0(0), b1(0),
          b2(0), b3(0), block(0xFFFFFFFFu) {}
    __device__ __forceinline__ uint32_t next_u32() {
        const uint32_t blk = n >> 2;
        if (blk != block) {
            uint32_t o[4];
            philox4x32_10(pixel, sample, blk, 0u, k0, k1, o);
            b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
            block = blk;
        }
        const uint32_t w = n & 3u;
        ++n;
        return w == 0 ? b0 : (w == 1 ? b1 : (w == 2 ? b2 : b3));
    }
    // rand 0.8.5 UniformFloat<f32>: 23 mantissa bits -> [1,2) - 1
    __device__ __forceinline__ float v01() { return __uint_as_float((next_u32() >> 9) | 0x3f800000u) - 1.0f; }
    // gen_range(low..high) — UniformFloat::sample_single
    __device__ __forceinline__ float gen_range(float low, float high) {
        const float scale = high - low;
        while (true) {
            const float res = v01() * scale + low;
            if (res < high) return res;
        }
    }
    // Uniform::new(-1.0, 1.0).sample
    __device__ __forceinline__ float uniform_m1_1() { return v01() * 2.0f + -1.0f; }
    __device__ __forceinline__ static float m1_1(uint32_t w) {  // uniform_m1_1 of one word
        return (__uint_as_float((w >> 9) | 0x3f800000u) - 1.0f) * 2.0f + -1.0f;
    }
    // The first accepted pair of rand_distr's rejection loops — loop { x1, x2 = Uniform(-1, 1); accept on x1^2 + x2^2 } —
    // with the acceptance INCLUSIVE (UnitDisc: <= 1) or not (UnitSphere, UnitCircle: < 1). A warp that loops until its
    // unluckiest lane is accepted runs ~3 iterations at a third of its lanes, each with a lazily computed Philox block
    // (40 % of k_raygen's instructions with a thin lens), so the first three candidates — the words n .. n + 5, which
    // live in the current block and the next — are evaluated up front at full width and the first accepted one is picked
    // without a branch; only a lane that rejects all three (1 %) enters the loop. Same draws, same count consumed.
    // UP_FRONT is the thin lens of k_raygen only (+1.1 % on config 1): in the shade kernels, which are not bound by
    // issue slots, the extra Philox block of every scatter costs more than the loop (-1.2 % on config 3).
    template <bool INCLUSIVE, bool UP_FRONT>
    __device__ __forceinline__ void accepted_pair(float& x1, float& x2, float& sum) {
        if (UP_FRONT && (n & 1u) == 0u) {
            const uint32_t blk = n >> 2;
            if (blk != block) {
                uint32_t o[4];
                philox4x32_10(pixel, sample, blk, 0u, k0, k1, o);
                b0 = o[0]; b1 = o[1]; b2 = o[2]; b3 = o[3];
                block = blk;
            }
            uint32_t c[4];
            philox4x32_10(pixel, sample, blk + 1u, 0u, k0, k1, c);
            const bool upper = (n & 2u) != 0u;  // the first pair is words 2, 3 of the current block
            const uint32_t w0 = upper ? b2 : b0, w1 = upper ? b3 : b1, w2 = upper ? c[0] : b2, w3 = upper ? c[1] : b3,
                           w4 = upper ? c[2] : c[0], w5 = upper ? c[3] : c[1];
            const float p1 = m1_1(w0), p2 = m1_1(w1), q1 = m1_1(w2), q2 = m1_1(w3), r1 = m1_1(w4), r2 = m1_1(w5);
            const float sp = p1 * p1 + p2 * p2, sq = q1 * q1 + q2 * q2, sr = r1 * r1 + r2 * r2;
            const bool okp = INCLUSIVE ? sp <= 1.0f : sp < 1.0f, okq = INCLUSIVE ? sq <= 1.0f : sq < 1.0f,
                       okr = INCLUSIVE ? sr <= 1.0f : sr < 1.0f;
            n += okp ? 2u : (okq ? 4u : 6u);
            if ((n >> 2) != blk) {  // the next draw is in the block just computed (or the one after it: computed when drawn)
                b0 = c[0]; b1 = c[1]; b2 = c[2]; b3 = c[3];
                block = blk + 1u;
            }
            x1 = okp ? p1 : (okq ? q1 : r1);
            x2 = okp ? p2 : (okq ? q2 : r2);
            sum = okp ? sp : (okq ? sq : sr);
            if (okp || okq || okr) return;
        }
        while (true) {
            x1 = uniform_m1_1();
            x2 = uniform_m1_1();
            sum = x1 * x1 + x2 * x2;
            if (INCLUSIVE ? sum <= 1.0f : sum < 1.0f) return;
        }
    }
    // rand_distr 0.4.3 UnitSphere (Marsaglia 1972)
    __device__ __forceinline__ f3 unit_sphere() {
        float x1, x2, sum;
        accepted_pair<false, false>(x1, x2, sum);
        const float factor = 2.0f * sqrtf(1.0f - sum);
        return f3{x1 * factor, x2 * factor, 1.0f - 2.0f * sum};
    }
    // rand 0.8.5 Standard f32: 24 random bits * 2^-24
    __device__ __forceinline__ float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    // BlockRng::next_u64: two consecutive words, low word first
    __device__ __forceinline__ unsigned long long next_u64() {
        const unsigned long long lo = next_u32();
        const unsigned long long hi = next_u32();
        return (hi << 32) | lo;
    }
    // rng.gen_bool(p as f64) = Bernoulli: p_int = (p * 2^64) as u64; p == 1 is always true without a draw
    __device__ __forceinline__ bool gen_bool(float pf) {
        const double p = (double)pf;
        if (p >= 1.0) return true;
        const unsigned long long p_int = p > 0.0 ? __double2ull_rz(p * 18446744073709551616.0) : 0ull;
        return next_u64() < p_int;
    }
    // rand_distr 0.4.3 UnitCircle
    __device__ __forceinline__ f2 unit_circle() {
        float x1, x2, sum;
        accepted_pair<false, false>(x1, x2, sum);
        const float diff = x1 * x1 - x2 * x2;
        return f2{diff / sum, 2.0f * x1 * x2 / sum};
    }
    // rand_distr 0.4.3 UnitDisc
    __device__ __forceinline__ f2 unit_disc() {
        float x1, x2, sum;
        accepted_pair<true, true>(x1, x2, sum);
        return f2{x1, x2};
    }
};

// util/math.rs
__device__ __forceinline__ bool near_zero(f3 v) {
    const float EPS = 1.0e-8f;
    return fabsf(v.x) < EPS && fabsf(v.y) < EPS && fabsf(v.z) < EPS;
}
__device__ __forceinline__ f3 reflect(f3 v, f3 n) { return v - 2.0f * dot(v, n) * n; }
__device__ __forceinline__ f3 refract(f3 uv, f3 n, float etai_over_etat) {
    const float cos_theta = fminf(dot(n, -uv), 1.0f);
    const f3 out_perp = etai_over_etat * (uv + cos_theta * n);
    const f3 out_parallel = -sqrtf(fabsf(1.0f - magnitude2(out_perp))) * n;
    return out_perp + out_parallel;
}
__device__ __forceinline__ f3 lerp3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }

}  // namespace vr
