// rng_pairs_check.cpp — test infrastructure: Rng::accepted_pair (voidray_b200/csrc/device_math.cuh: the first three
// candidates of rand_distr's rejection loops evaluated up front) against the plain loop over Rng::uniform_m1_1, for
// UnitSphere, UnitCircle and UnitDisc: same values bit for bit, same number of draws consumed, same stream afterwards —
// from every starting draw count 0 .. 11 (both pair positions of a block, odd counts, a cached or a cold block).
//   g++ -O2 -std=c++17 -ffp-contract=off -DVR_HOST_SHIM -Itests/c -Ivoidray_b200/csrc tests/c/rng_pairs_check.cpp -o rng_pairs_check
#include <cstdio>
#include <cstring>

#include "host_shim.h"
#include "device_math.cuh"

using namespace vr;

static bool same(float a, float b) { return std::memcmp(&a, &b, 4) == 0; }

int main() {
    unsigned long long checked = 0, wrong = 0, fallbacks = 0;
    for (uint32_t pixel = 0; pixel < 3000; ++pixel)
        for (uint32_t sample = 0; sample < 8; ++sample)
            for (uint32_t n0 = 0; n0 < 12; ++n0)
                for (int warm = 0; warm < 2; ++warm)
                    for (int kind = 0; kind < 3; ++kind) {
                        Rng a(0x5EED0001ull + pixel * 77ull, pixel, sample, n0), b(0x5EED0001ull + pixel * 77ull, pixel, sample, n0);
                        if (warm && n0 > 0) {  // a block already cached by an earlier draw
                            a.n = b.n = n0 - 1;
                            a.next_u32();
                            b.next_u32();
                        }
                        float r[3] = {0, 0, 0}, w[3] = {0, 0, 0};
                        // the plain loops (rand_distr 0.4.3)
                        float x1, x2, sum;
                        while (true) {
                            x1 = b.uniform_m1_1();
                            x2 = b.uniform_m1_1();
                            sum = x1 * x1 + x2 * x2;
                            if (kind == 2 ? sum <= 1.0f : sum < 1.0f) break;
                        }
                        if (b.n - n0 > 6) ++fallbacks;
                        if (kind == 0) {
                            const float factor = 2.0f * sqrtf(1.0f - sum);
                            w[0] = x1 * factor; w[1] = x2 * factor; w[2] = 1.0f - 2.0f * sum;
                            const f3 v = a.unit_sphere();
                            r[0] = v.x; r[1] = v.y; r[2] = v.z;
                        } else if (kind == 1) {
                            const float diff = x1 * x1 - x2 * x2;
                            w[0] = diff / sum; w[1] = 2.0f * x1 * x2 / sum;
                            const f2 v = a.unit_circle();
                            r[0] = v.x; r[1] = v.y;
                        } else {
                            w[0] = x1; w[1] = x2;
                            const f2 v = a.unit_disc();
                            r[0] = v.x; r[1] = v.y;
                        }
                        bool ok = a.n == b.n && same(r[0], w[0]) && same(r[1], w[1]) && same(r[2], w[2]);
                        for (int k = 0; k < 9 && ok; ++k) ok = a.next_u32() == b.next_u32();  // the stream goes on identically
                        ++checked;
                        if (!ok) {
                            if (wrong < 10) std::printf("pixel %u sample %u n0 %u warm %d kind %d: n %u vs %u\n", pixel, sample, n0, warm, kind, a.n, b.n);
                            ++wrong;
                        }
                    }
    std::printf("%llu streams checked (%llu past the three up-front candidates), %llu wrong\n", checked, fallbacks, wrong);
    return wrong || fallbacks == 0 ? 1 : 0;
}
