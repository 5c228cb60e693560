/* The VR_TEX8 experiment widens 8-bit texels on the device as q = v * (1/255); q += fma(-q, 255, v) * (1/255)
 * (kernels.cu unorm8). This checks on the host, for every v, that the result is bit-identical to the correctly rounded
 * (float)v / 255.0f of to_rgb32f. Compile with -ffp-contract=off. */
#include <math.h>
#include <stdio.h>
#include <string.h>

int main(void) {
    const float r = 1.0f / 255.0f;
    int bad = 0, v;
    for (v = 0; v < 256; ++v) {
        const float x = (float)v, want = x / 255.0f;
        const float q = x * r;
        const float got = fmaf(fmaf(-q, 255.0f, x), r, q);
        if (memcmp(&got, &want, 4) != 0) ++bad;
    }
    printf("unorm8: %d of 256 values differ\n", bad);
    return bad != 0;
}
