// fastdiv_check.cpp — test infrastructure: make_fast_div / the kernels' fast_div (voidray_b200/csrc/kernels.cuh, kernels.cu)
// against the plain division. Every divisor in [1, 70000) plus image-sized ones (pixel counts, tiles per row) over
// dividends at the boundaries of each quotient and a pseudo-random sweep below 2^31.
//   g++ -O2 -std=c++20 -pthread -DVR_HOST_SHIM -DVR_HOST_SIMT -Itests/c -Ivoidray_b200/csrc tests/c/fastdiv_check.cpp -o fastdiv_check
#include <cstdio>
#include <vector>

#include "kernels.cuh"

static inline uint32_t fast_div(uint32_t x, const vr::FastDiv& f) {  // as in kernels.cu
    if (f.m == 0u) return f.d <= 1u ? x : x / f.d;
    return __umulhi(x, f.m) >> f.s;
}

int main() {
    std::vector<uint32_t> divisors;
    for (uint32_t d = 1; d < 70000u; ++d) divisors.push_back(d);
    for (uint32_t d : {480000u, 2073600u, 8294400u, 33554432u, 67108863u, 1000003u, 2147483647u, 1073741824u, 1073741825u})
        divisors.push_back(d);
    unsigned long long checked = 0, wrong = 0;
    uint64_t lcg = 0x9E3779B97F4A7C15ull;
    for (uint32_t d : divisors) {
        const vr::FastDiv f = vr::make_fast_div(d);
        auto check = [&](uint64_t x64) {
            if (x64 >= (1ull << 31)) return;
            const uint32_t x = (uint32_t)x64;
            ++checked;
            if (fast_div(x, f) != x / d) {
                if (wrong < 10) std::printf("d=%u x=%u: %u != %u\n", d, x, fast_div(x, f), x / d);
                ++wrong;
            }
        };
        check(0);
        check((1ull << 31) - 1);
        for (int k = 0; k < 64; ++k) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const uint64_t x = (lcg >> 33);
            check(x);
            const uint64_t q = x / d;
            check(q * d);
            check(q * d + d - 1);
            if (q * d) check(q * d - 1);
        }
    }
    std::printf("%llu divisions checked, %llu wrong\n", checked, wrong);
    return wrong ? 1 : 0;
}
