// trav_host.cpp — test infrastructure: the closest-hit traversal of the CUDA kernels (voidray_b200/csrc/traversal.cuh,
// the very source kernels.cu includes) compiled for the CPU through tests/c/host_shim.h, run over the scene the
// library's own flattener produces. tests/test_traversal_host.py compares its hits with the oracle bit for bit —
// the k_trace_rays gate without a GPU, for the shipped layout and for every experiment variant
// (-DVR_BVH4, -DVR_TRI48, -DVR_SMEM_STACK=n: the same flags for this file and for scene_build.cpp).
//
//   g++ -O2 -std=c++17 -pthread -ffp-contract=off -DVR_HOST_SHIM -Itests/c -Ivoidray_b200/csrc -x c++
//       voidray_b200/csrc/scene_build.cpp tests/c/trav_host.cpp -o trav_host
//   trav_host rays.bin out.bin [obj <path> | sphere cx cy cz r | plane h]...
// rays.bin: n x (origin xyz, direction xyz) f32; out.bin: n x (surface u32, prim u32, t f32), k_trace_rays semantics.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "traversal.cuh"
#include "scene_build.h"

using namespace vr;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    HostScene sc;
    MaterialRec mat{};
    mat.albedo_tex = -1;
    mat.normal_tex = -1;
    sc.materials.push_back(mat);
    std::string err;
    for (int i = 3; i < argc; ++i) {
        const std::string kind = argv[i];
        HostSurface sf;
        if (kind == "obj" && i + 1 < argc) {
            HostMesh m;
            if (!load_obj_file(argv[++i], m, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
            sf.kind = 0;
            sf.mesh = (uint32_t)sc.meshes.size();
            sc.meshes.push_back(std::move(m));
        } else if (kind == "sphere" && i + 4 < argc) {
            sf.kind = 1;
            for (int a = 0; a < 3; ++a) sf.center[a] = (float)atof(argv[++i]);
            sf.radius_or_height = (float)atof(argv[++i]);
        } else if (kind == "plane" && i + 1 < argc) {
            sf.kind = 2;
            sf.radius_or_height = (float)atof(argv[++i]);
        } else {
            return 2;
        }
        sc.surfaces.push_back(sf);
        sc.objects.push_back(HostObject{(uint32_t)sc.surfaces.size() - 1, 0});
    }
    const float eye[3] = {0.0f, 0.0f, -1.0f}, dir[3] = {0.0f, 0.0f, 1.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    for (int a = 0; a < 3; ++a) { sc.camera.eye[a] = eye[a]; sc.camera.direction[a] = dir[a]; sc.camera.up[a] = up[a]; }
    sc.camera.fov = 0.6f;
    sc.camera.has_dof = 0;
    FlatScene flat;
    if (!flatten_scene(sc, flat, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }

    DeviceScene ds{};
    ds.nodes = flat.nodes.data();
    ds.tri_isect = flat.tri_isect.data();
    ds.tri_shade = flat.tri_shade.data();
    ds.tri_surface = flat.tri_surface.data();
    ds.tri_prim = flat.tri_prim.data();
    ds.analytics = flat.analytics.data();
    ds.n_analytics = (uint32_t)flat.analytics.size();
    ds.n_tris = flat.n_tris;
    ds.scene_tree = flat.scene_tree.data();
    ds.surface_node = flat.surface_node.data();
    ds.n_scene_nodes = (uint32_t)flat.scene_tree.size();
    ds.n_surfaces = (uint32_t)flat.surface_node.size();
    for (int a = 0; a < 3; ++a) { ds.grid_min[a] = flat.grid_min[a]; ds.grid_extent[a] = flat.grid_extent[a]; }

    FILE* in = std::fopen(argv[1], "rb");
    if (!in) return 1;
    std::vector<float> rays;
    float buf[6 * 1024];
    size_t got;
    while ((got = std::fread(buf, sizeof(float), 6 * 1024, in)) > 0) rays.insert(rays.end(), buf, buf + got);
    std::fclose(in);
    const size_t n = rays.size() / 6;

    struct Out { uint32_t surface, prim; float t; };
    std::vector<Out> out(n);
    // one stack column, stride 1 (the kernel: one column per thread, stride = block size)
    std::vector<int> stack(STACK_DEPTH + 8, 0);
    int max_sp_seen = 0;
    (void)max_sp_seen;
    for (size_t i = 0; i < n; ++i) {
        const f3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        const f3 d = normalize(mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));  // Ray::new, as k_trace_rays
        const HitResult h = closest_hit(ds, o, d, stack.data(), 1);
        if (h.prim < 0) out[i] = Out{0xFFFFFFFFu, 0xFFFFFFFFu, INFINITY};
        else if ((uint32_t)h.prim < ds.n_tris) out[i] = Out{ds.tri_surface[h.prim], ds.tri_prim[h.prim], h.t};
        else out[i] = Out{ds.analytics[h.prim - ds.n_tris].surface, 0xFFFFFFFFu, h.t};
    }
    FILE* of = std::fopen(argv[2], "wb");
    if (!of) return 1;
    std::fwrite(out.data(), sizeof(Out), n, of);
    std::fclose(of);
    std::printf("%zu rays, %u triangles, %zu nodes of %d quads, %zu analytic surfaces\n", n, flat.n_tris,
                flat.nodes.size() / DEVICE_NODE_QUADS, DEVICE_NODE_QUADS, flat.analytics.size());
    return 0;
}
