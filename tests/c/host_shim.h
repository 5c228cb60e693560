// host_shim.h — test infrastructure only: stands in for the CUDA intrinsics that voidray_b200/csrc/device_math.cuh
// and traversal.cuh use, so that tests/c/trav_host.cpp can compile the kernels' own traversal source with g++
// (-DVR_HOST_SHIM -ffp-contract=off: like nvcc -fmad=false, every a * b + c is two roundings; FMA only where the
// source says __fmaf_rn). Never part of the library.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __align__(n) alignas(n)

struct alignas(16) float4 {
    float x, y, z, w;
};
struct uchar4 {
    unsigned char x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <class T>
inline T __ldg(const T* p) { return *p; }

inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
// PRMT, default mode: byte i of the result = byte (selector nibble i & 7) of {y, x}; nibble bit 3 replicates the sign
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t both = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xFu;
        uint32_t b = (uint32_t)(both >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) b = (b & 0x80u) ? 0xFFu : 0x00u;
        r |= b << (8 * i);
    }
    return r;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t __float2uint_rz(float x) {  // saturating, NaN -> 0
    if (!(x > 0.0f)) return 0u;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
inline unsigned long long __double2ull_rz(double x) {
    if (!(x > 0.0)) return 0ull;
    if (x >= 18446744073709551616.0) return ~0ull;
    return (unsigned long long)x;
}
