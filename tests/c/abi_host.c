/* Strict-C99 consumer of include/voidray_cuda.h: proves the header is plain C and exercises the entry points that need
 * no device. Built and run by tests/test_c_header.py. */
#include <stdio.h>
#include <stdlib.h>
#include "voidray_cuda.h"
int main(int argc, char** argv) {
    uint32_t w = 0, h = 0;
    float* rgb = NULL;
    float eye[3] = {0, 0, 5}, center[3] = {0, 0, 0}, up[3] = {0, 1, 0}, dir[3], up2[3];
    float boxes[12] = {0, 0, 0, 1, 1, 1, 2, 0, 0, 3, 1, 1};
    uint32_t order[2];
    vr_context* ctx = NULL;
    if (vr_abi_version() != 2) return 1;
    if (vr_camera_look_at(eye, center, up, dir, up2) != VR_OK) return 2;
    if (vr_debug_reference_leaf_order(boxes, 2, order) != VR_OK || order[0] != 0 || order[1] != 1) return 3;
    if (argc > 1) {
        if (vr_image_load_rgb32f(argv[1], &w, &h, &rgb) != VR_OK) { printf("%s\n", vr_last_error()); return 4; }
        printf("%u x %u first texel %.6f %.6f %.6f\n", w, h, rgb[0], rgb[1], rgb[2]);
        vr_image_free(rgb);
    }
    if (vr_context_create(0, NULL, &ctx) == VR_OK) vr_context_destroy(ctx);
    else printf("no device: %s\n", vr_last_error());
    printf("dir %.3f %.3f %.3f sizeof(settings) %zu sizeof(stats) %zu sizeof(material) %zu\n", dir[0], dir[1], dir[2], sizeof(vr_render_settings), sizeof(vr_stats), sizeof(vr_material_desc));
    return 0;
}
