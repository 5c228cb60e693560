// wavefront_host.cpp — test infrastructure: the CUDA kernels of voidray_b200/csrc/kernels.cu themselves (k_raygen,
// k_trace, k_shade, k_accumulate — the very source nvcc compiles) built for the CPU through tests/c/host_shim.h with
// -DVR_HOST_SHIM -DVR_HOST_SIMT: one OS thread per CUDA thread, warp-level intrinsics over a barrier per warp, real
// atomics. Runs one wavefront batch of a small frame the way run_wavefront / vr_render_accumulate (csrc/abi.cu) drive
// the device, and writes the accumulation buffer and the per-ray closest hits of depth 0 and 1.
// tests/test_wavefront_host.py compares them with the oracle: the kernels' warp-level control flow (ray replacement,
// vote stepping, queue compaction, the -DVR_TRACE_CHUNK claims, the -DVR_BVH4 step) checked without a GPU.
//
//   g++ -O2 -std=c++20 -pthread -ffp-contract=off -DVR_HOST_SHIM -DVR_HOST_SIMT -Itests/c -Ivoidray_b200/csrc -x c++
//       voidray_b200/csrc/scene_build.cpp tests/c/wavefront_host.cpp -o wavefront_host
//   wavefront_host <obj> <w> <h> <spp> <max_bounces> <seed> <eye xyz> <center xyz> <fov> <env rgb> <albedo rgb> <out.bin>
//   wavefront_host <scene file of tests/scene_file.py> <w> <h> <spp> <max_bounces> <seed> <firefly_clamp> <render_mode>
//                  <pixel_mapping> <out.bin>            (any scene the host API can describe)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "scene_build.h"
#include "../../voidray_b200/csrc/kernels.cu"
#include "scene_file.h"

using namespace vr;

int main(int argc, char** argv) {
    if (argc < 11) return 2;
    int a = 1;
    const std::string first = argv[a++];
    const bool scene_mode = first.size() > 8 && first.compare(first.size() - 8, 8, ".vrscene") == 0;
    if (!scene_mode && argc < 20) return 2;
    const uint32_t w = (uint32_t)atoi(argv[a++]), h = (uint32_t)atoi(argv[a++]), spp = (uint32_t)atoi(argv[a++]);
    const uint32_t max_bounces = (uint32_t)atoi(argv[a++]);
    const uint64_t seed = strtoull(argv[a++], nullptr, 0);
    float firefly_clamp = 3.0f;
    int32_t render_mode = 0, pixel_mapping = 0;
    HostScene sc;
    std::string err;
    if (scene_mode) {
        firefly_clamp = (float)atof(argv[a++]);
        render_mode = atoi(argv[a++]);
        pixel_mapping = atoi(argv[a++]);
        if (!vr_test::read_scene(first.c_str(), sc)) { std::fprintf(stderr, "cannot read %s\n", first.c_str()); return 1; }
    } else {
        float eye[3], center[3], env[3], albedo[3];
        for (float& x : eye) x = (float)atof(argv[a++]);
        for (float& x : center) x = (float)atof(argv[a++]);
        const float fov = (float)atof(argv[a++]);
        for (float& x : env) x = (float)atof(argv[a++]);
        for (float& x : albedo) x = (float)atof(argv[a++]);
        MaterialRec mat{};
        mat.kind = 0;  // Lambertian
        for (int k = 0; k < 3; ++k) mat.color[k] = albedo[k];
        mat.albedo_tex = -1;
        mat.normal_tex = -1;
        sc.materials.push_back(mat);
        HostMesh m;
        if (!load_obj_file(first.c_str(), m, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
        sc.meshes.push_back(std::move(m));
        HostSurface sf;
        sf.kind = 0;
        sf.mesh = 0;
        sc.surfaces.push_back(sf);
        sc.objects.push_back(HostObject{0, 0});
        const float up[3] = {0.0f, 1.0f, 0.0f};
        for (int k = 0; k < 3; ++k) sc.camera.eye[k] = eye[k];
        camera_look_at(eye, center, up, sc.camera.direction, sc.camera.up);
        sc.camera.fov = fov;
        sc.camera.has_dof = 0;
        sc.env_kind = 1;
        for (int k = 0; k < 3; ++k) sc.env_color[k] = env[k];
    }
    const char* out_path = argv[a++];
    vr_test::HostDeviceScene hds;
    if (!hds.build(sc, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const DeviceScene& ds = hds.ds;

    const uint32_t n_pixels = w * h, n_paths = n_pixels * spp;
    std::vector<float4> ray_o(n_paths), ray_d(n_paths), ray_o1(n_paths), ray_d1(n_paths), hit(n_paths), radiance(n_paths),
        att((size_t)max_bounces * n_paths);
    std::vector<uint32_t> q0(n_paths), q1(n_paths), counts(2 * (max_bounces + 2) + 1, 0u);
    std::vector<float4> miss(n_paths);
    unsigned long long segments = 0, culled = 0;
    Wavefront wf{};
    wf.ray_o[0] = ray_o.data();
    wf.ray_d[0] = ray_d.data();
    wf.ray_o[1] = ray_o1.data();
    wf.ray_d[1] = ray_d1.data();
    wf.hit = hit.data();
    wf.att = att.data();
    wf.radiance = radiance.data();
    wf.queue[0] = q0.data();
    wf.queue[1] = q1.data();
    wf.counts = counts.data();
    wf.cursors = counts.data() + (max_bounces + 2);
    wf.miss = miss.data();
    wf.miss_count = counts.data() + 2 * (max_bounces + 2);
    wf.segments = &segments;
    wf.culled = &culled;
    wf.capacity = n_paths;
    const PathSource src = make_path_source(nullptr, nullptr, w, h, 0);
    FrameParams fp{};
    fp.width = w;
    fp.height = h;
    fp.pixel_mapping = pixel_mapping;
    fp.max_bounces = max_bounces;
    fp.firefly_clamp = firefly_clamp;
    fp.render_mode = render_mode;
    fp.integrator = 0;
    fp.seed = seed;

    // what run_wavefront does (csrc/abi.cu), with small grids: 2 blocks of ray generation, 2 persistent trace blocks
    // (8 warps racing for the queue), 2 shade blocks
    std::vector<float4> hits_depth0, hits_depth1;
    std::vector<uint32_t> queue_depth1;
    // every camera ray in slot order first (cull = 0, what the gate kernels use): the reference for the compacted run
    vr_host_launch(2, 256, [&] { k_raygen(ds, wf, src, fp, n_paths, spp > 1 ? 2u : 1u, 0); });
    const std::vector<float4> all_o = ray_o, all_d = ray_d;
    std::fill(counts.begin(), counts.end(), 0u);
    segments = 0;
    culled = 0;
    // the shipped path: rays that miss the scene's bounds are finished by k_raygen, the rest compacted into queue 0
    vr_host_launch(2, 256, [&] { k_raygen(ds, wf, src, fp, n_paths, spp > 1 ? 2u : 1u, 1); });
    const uint32_t n_queued = counts[0];
    const std::vector<uint32_t> queue0(q0.begin(), q0.begin() + n_queued);
    const std::vector<float4> rays0_o(ray_o.begin(), ray_o.begin() + n_queued), rays0_d(ray_d.begin(), ray_d.begin() + n_queued);
    std::vector<float4> rays1_o, rays1_d;
    for (uint32_t depth = 0; depth < max_bounces; ++depth) {
        if (depth == 1) { rays1_o = ray_o1; rays1_d = ray_d1; queue_depth1.assign(q1.begin(), q1.begin() + counts[1]); }
        vr_host_launch(2, TRACE_THREADS, [&] { k_trace(ds, wf, depth, REFILL_THRESHOLD); });
        if (depth == 0) hits_depth0 = hit;
        if (depth == 1) hits_depth1 = hit;
        if (depth == 0) {
            if (ds.has_microfacet) vr_host_launch(2, SHADE_THREADS, [&] { k_shade_first<false, true>(ds, wf, src, fp); });
            else vr_host_launch(2, SHADE_THREADS, [&] { k_shade_first<false, false>(ds, wf, src, fp); });
        } else if (ds.has_microfacet) vr_host_launch(2, SHADE_THREADS, [&] { k_shade<false, true>(ds, wf, src, fp, depth); });
        else vr_host_launch(2, SHADE_THREADS, [&] { k_shade<false, false>(ds, wf, src, fp, depth); });
    }
    if (max_bounces > 1) vr_host_launch(2, 256, [&] { k_miss(ds, wf, fp.firefly_clamp); });
    std::vector<float4> partial(n_pixels, float4{0, 0, 0, 0}), accum(n_pixels, float4{0, 0, 0, 0});
    vr_host_launch(1, 256, [&] { k_accumulate(wf, partial.data(), accum.data(), w, h, spp, 1, 1.0f / (float)spp, 1.0f); });

    // every ray of depth 0 and of depth 1 against the single-ray traversal (closest_hit, the gate kernels' path); the
    // depth-0 queue holds each slot at most once with the ray k_raygen generates for it, and every slot it does not
    // hold is a certain miss
    std::vector<int> stack(STACK_DEPTH + 8);
    size_t wrong = 0, checked = 0;
    auto check = [&](const std::vector<float4>& ro, const std::vector<float4>& rd, const std::vector<float4>& got, uint32_t slot) {
        const HitResult hr = closest_hit(ds, xyz(ro[slot]), xyz(rd[slot]), stack.data(), 1);
        const float4 g = got[slot];
        ++checked;
        if (__float_as_uint(g.x) != __float_as_uint(hr.t) || __float_as_int(g.y) != hr.prim || __float_as_uint(g.z) != __float_as_uint(hr.u) ||
            __float_as_uint(g.w) != __float_as_uint(hr.v))
            ++wrong;
    };
    std::vector<char> queued(n_paths, 0);
    for (uint32_t i = 0; i < n_queued; ++i) {
        const uint32_t slot = queue0[i];
        ++checked;
        if (slot >= n_paths || queued[slot] || std::memcmp(&rays0_o[i], &all_o[slot], 16) != 0 || std::memcmp(&rays0_d[i], &all_d[slot], 16) != 0) {
            ++wrong;
            continue;
        }
        queued[slot] = 1;
        check(rays0_o, rays0_d, hits_depth0, i);  // queue order
    }
    unsigned long long n_culled = 0;
    for (uint32_t slot = 0; slot < n_paths; ++slot) {
        if (queued[slot]) continue;
        ++checked;
        ++n_culled;
        if (closest_hit(ds, xyz(all_o[slot]), xyz(all_d[slot]), stack.data(), 1).prim >= 0) ++wrong;  // culled, but it hits
    }
    if (n_culled != culled) ++wrong;
    for (uint32_t i = 0; i < (uint32_t)queue_depth1.size(); ++i) check(rays1_o, rays1_d, hits_depth1, i);  // queue order
    std::printf("%u paths, queue lengths", n_paths);
    for (uint32_t d = 0; d <= max_bounces; ++d) std::printf(" %u", counts[d]);
    std::printf(", %llu segments, %zu of %zu wavefront hits differ from the single-ray traversal\n", segments, wrong, checked);
    FILE* of = std::fopen(out_path, "wb");
    if (!of) return 1;
    std::fwrite(accum.data(), sizeof(float4), n_pixels, of);
    std::fclose(of);
    return wrong ? 1 : 0;
}
