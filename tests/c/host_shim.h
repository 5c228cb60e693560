// host_shim.h — test infrastructure only: stands in for the CUDA intrinsics that voidray_b200/csrc/device_math.cuh
// and traversal.cuh use, so that tests/c/trav_host.cpp can compile the kernels' own traversal source with g++
// (-DVR_HOST_SHIM -ffp-contract=off: like nvcc -fmad=false, every a * b + c is two roundings; FMA only where the
// source says __fmaf_rn). Never part of the library.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

using std::isfinite;
using std::isnan;

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

// ---- optional: whole kernels on the CPU (-DVR_HOST_SIMT, C++20) -----------------------------------------------------
// A block runs as one OS thread per CUDA thread; the 32 threads of a warp meet at a barrier inside every warp-level
// intrinsic, so __ballot_sync / __shfl_sync / __any_sync have their CUDA meaning as long as the kernel calls them under
// warp-uniform control flow with a full mask (k_trace and k_shade do). Blocks run one after the other, so a __shared__
// array is a function-local static. Global atomics are real atomics. Used by tests/c/ktrace_host.cpp.
#ifdef VR_HOST_SIMT
#include <barrier>
#include <thread>
#include <vector>

#define __shared__ static
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
typedef void* cudaStream_t;

struct vr_uint3 {
    unsigned x, y, z;
};
inline thread_local vr_uint3 threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
struct vr_warp_ctx {
    std::barrier<> bar{32};
    uint32_t vals[32];
};
inline thread_local vr_warp_ctx* vr_warp = nullptr;
inline thread_local std::barrier<>* vr_block = nullptr;
inline void __syncthreads() { vr_block->arrive_and_wait(); }  // block-uniform control flow only, as on the device

inline unsigned __ballot_sync(unsigned, bool p) {
    vr_warp_ctx& w = *vr_warp;
    w.vals[threadIdx.x & 31u] = p ? 1u : 0u;
    w.bar.arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= w.vals[i] << i;
    w.bar.arrive_and_wait();
    return m;
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) {
    vr_warp_ctx& w = *vr_warp;
    w.vals[threadIdx.x & 31u] = v;
    w.bar.arrive_and_wait();
    const uint32_t r = w.vals[src & 31];
    w.bar.arrive_and_wait();
    return r;
}
inline uint32_t __shfl_xor_sync(unsigned mask, uint32_t v, int lane_mask) {
    return __shfl_sync(mask, v, (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask));
}
inline void __syncwarp() {
    vr_warp->bar.arrive_and_wait();
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// kernel<<<grid, block>>>(args...) : blocks one after the other, the threads of a block concurrently
template <class Kernel>
inline void vr_host_launch(unsigned grid, unsigned block, Kernel&& kernel) {
    std::vector<vr_warp_ctx> warps((block + 31) / 32);
    for (unsigned b = 0; b < grid; ++b) {
        std::barrier<> block_bar(block);
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t, b] {
                threadIdx = vr_uint3{t, 0, 0};
                blockIdx = vr_uint3{b, 0, 0};
                blockDim = vr_uint3{block, 1, 1};
                gridDim = vr_uint3{grid, 1, 1};
                vr_warp = &warps[t / 32];
                vr_block = &block_bar;
                kernel();
            });
        for (std::thread& th : threads) th.join();
    }
}
#endif
