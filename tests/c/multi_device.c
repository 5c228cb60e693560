/* In-process device group through include/voidray_cuda.h alone (strict C99, no Python, no torch):
 *   multi_device <mesh.obj> <dev,dev,...>
 * renders the same scene on one device and on the device group (vr_context_create_multi: commit replicates the scene,
 * vr_render_accumulate shards the samples, vr_render_read_accum / vr_render_resolve sum the shards over peer memory)
 * and compares the two accumulation buffers (<= 2e-6) and resolved images. Run by tests/test_gpu_multi.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "voidray_cuda.h"

#define W 160u
#define H 120u
#define SPP 16u
#define CHECK(call)                                                        \
    do {                                                                   \
        if ((call) != VR_OK) {                                             \
            printf("%s failed: %s\n", #call, vr_last_error());             \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static int build_scene(vr_context* ctx, const char* obj, vr_scene** out) {
    vr_scene* s = NULL;
    vr_material_desc m;
    uint32_t surface = 0, material = 0, object = 0, sphere = 0, metal = 0;
    float eye[3] = {0.2f, 2.8f, -10.5f}, center[3] = {0.2f, 0.8f, -0.5f}, up[3] = {0.0f, 1.0f, 0.0f};
    float sky[3] = {0.7f, 0.8f, 0.9f}, c[3] = {1.2f, 0.4f, 0.0f};
    CHECK(vr_scene_create(ctx, &s));
    CHECK(vr_scene_add_mesh_from_obj_file(s, obj, &surface, NULL, NULL));
    memset(&m, 0, sizeof m);
    m.kind = VR_MAT_LAMBERTIAN;
    m.color[0] = 0.6f; m.color[1] = 0.5f; m.color[2] = 0.4f;
    m.albedo_tex = -1; m.normal_tex = -1;
    CHECK(vr_scene_add_material(s, &m, &material));
    CHECK(vr_scene_add_object(s, material, surface, &object));
    CHECK(vr_scene_add_sphere(s, c, 0.4f, &sphere));
    m.kind = VR_MAT_METAL;
    m.param = 0.1f;
    CHECK(vr_scene_add_material(s, &m, &metal));
    CHECK(vr_scene_add_object(s, metal, sphere, &object));
    CHECK(vr_scene_set_camera_look_at(s, eye, center, up, 0.17f));
    CHECK(vr_scene_set_environment_uniform(s, sky));
    CHECK(vr_scene_commit(s));
    *out = s;
    return 0;
}

static int render(vr_context* ctx, const char* obj, const uint32_t* calls, int n_calls, float* accum, float* resolved,
                  vr_stats* stats) {
    vr_scene* s = NULL;
    vr_render* r = NULL;
    vr_render_settings st;
    int i;
    if (build_scene(ctx, obj, &s)) return 1;
    memset(&st, 0, sizeof st);
    st.total_samples = SPP;
    st.max_bounces = 6;
    st.firefly_clamp = 3.0f;
    st.seed = 0x5EED0001u;
    CHECK(vr_render_begin(s, W, H, &st, &r));
    for (i = 0; i < n_calls; ++i) CHECK(vr_render_accumulate(r, calls[i]));
    CHECK(vr_render_read_accum(r, accum));
    CHECK(vr_render_resolve(r, 1.0f, 2.2f, 0.5f, 1, resolved));
    CHECK(vr_render_read_accum(r, accum)); /* the reduce is non-destructive: reading twice gives the same buffer */
    CHECK(vr_render_stats(r, stats));
    CHECK(vr_render_end(r));
    CHECK(vr_scene_destroy(s));
    return 0;
}

int main(int argc, char** argv) {
    int32_t ids[16];
    uint32_t n = 0, n_reported = 0;
    vr_context *one = NULL, *group = NULL;
    const size_t count = (size_t)W * H * 4;
    float *a = malloc(count * 4), *b = malloc(count * 4), *ra = malloc(count * 4), *rb = malloc(count * 4);
    const uint32_t one_call[1] = {SPP}, uneven[3] = {1, 5, SPP - 6}; /* 1 sample: fewer samples than devices */
    vr_stats sa, sb;
    double d_acc = 0.0, d_res = 0.0;
    size_t i;
    char* tok;
    if (argc < 3 || !a || !b || !ra || !rb) return 2;
    for (tok = strtok(argv[2], ","); tok && n < 16; tok = strtok(NULL, ",")) ids[n++] = atoi(tok);
    CHECK(vr_context_create(ids[0], NULL, &one));
    CHECK(vr_context_create_multi(ids, n, &group));
    CHECK(vr_context_device_count(group, &n_reported));
    if (n_reported != n) { printf("device count %u != %u\n", n_reported, n); return 1; }
    if (render(one, argv[1], one_call, 1, a, ra, &sa)) return 1;
    if (render(group, argv[1], one_call, 1, b, rb, &sb)) return 1;
    for (i = 0; i < count; ++i) {
        const double d = fabs((double)a[i] - (double)b[i]);
        if ((i & 3) == 3) { if (a[i] != b[i]) { printf("alpha differs: %g vs %g\n", a[i], b[i]); return 1; } }
        else if (d > d_acc) d_acc = d;
        if (fabs((double)ra[i] - (double)rb[i]) > d_res) d_res = fabs((double)ra[i] - (double)rb[i]);
    }
    printf("one call: max |accum diff| %.3g, max |resolved diff| %.3g, segments %llu vs %llu\n", d_acc, d_res,
           (unsigned long long)sa.ray_segments, (unsigned long long)sb.ray_segments);
    if (d_acc > 2e-6 || d_res > 2e-5 || sa.ray_segments != sb.ray_segments || sb.samples_done != SPP) return 1;
    /* progressive calls of uneven size (the 1-spp probe of RenderThread::one_shot gives some devices nothing) */
    if (render(one, argv[1], uneven, 3, a, ra, &sa)) return 1;
    if (render(group, argv[1], uneven, 3, b, rb, &sb)) return 1;
    d_acc = 0.0;
    for (i = 0; i < count; ++i) {
        const double d = fabs((double)a[i] - (double)b[i]);
        if ((i & 3) == 3) { if (a[i] != b[i] || a[i] != 3.0f) { printf("alpha differs: %g vs %g\n", a[i], b[i]); return 1; } }
        else if (d > d_acc) d_acc = d;
    }
    printf("three calls: max |accum diff| %.3g\n", d_acc);
    if (d_acc > 2e-6 || sa.ray_segments != sb.ray_segments) return 1;
    CHECK(vr_context_destroy(group));
    CHECK(vr_context_destroy(one));
    printf("ok %u devices\n", n);
    free(a); free(b); free(ra); free(rb);
    return 0;
}
