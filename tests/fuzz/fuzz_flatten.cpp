// Random degenerate scenes (NaN / infinite / duplicate vertices, empty meshes, analytic surfaces) through flatten_scene
// (csrc/scene_build.cpp), built with ASan + UBSan by tests/test_fuzz_host.py; checks that every triangle is reachable
// exactly once and the tree fits the traversal stack.
//   fuzz_flatten <seed> <scenes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include "scene_build.h"
using namespace vr;
int main(int argc, char** argv) {
    std::mt19937 rng(argc > 1 ? atoi(argv[1]) : 1);
    int iters = argc > 2 ? atoi(argv[2]) : 2000;
    const float specials[] = {0.0f, -0.0f, 1.0f, -1.0f, INFINITY, -INFINITY, NAN, 1e-30f, 1e30f, -1e30f, 3.4e38f, 1e-45f};
    for (int it = 0; it < iters; ++it) {
        HostScene sc;
        int n_surf = 1 + rng() % 4;
        MaterialRec m{}; m.albedo_tex = -1; m.normal_tex = -1; sc.materials.push_back(m);
        for (int s = 0; s < n_surf; ++s) {
            int kind = rng() % 5;
            HostSurface sf;
            if (kind <= 2) {
                HostMesh mesh;
                int nv = 1 + rng() % 60, nt = rng() % 200;
                int mode = rng() % 4;
                for (int v = 0; v < nv; ++v) {
                    for (int a = 0; a < 3; ++a) {
                        float x = mode == 0 ? std::uniform_real_distribution<float>(-5, 5)(rng) : mode == 1 ? (float)(rng() % 3) : mode == 2 ? specials[rng() % 12] : (rng() % 10 == 0 ? specials[rng() % 12] : std::uniform_real_distribution<float>(-5, 5)(rng));
                        mesh.pos.push_back(x);
                        mesh.nrm.push_back(mode == 2 ? specials[rng() % 12] : 0.5f);
                    }
                    mesh.uv.push_back(0.25f); mesh.uv.push_back(specials[rng() % 12]);
                }
                mesh.n_vertices = nv;
                for (int t = 0; t < 3 * nt; ++t) mesh.idx.push_back(rng() % nv);
                sc.meshes.push_back(mesh);
                sf.kind = 0; sf.mesh = (uint32_t)sc.meshes.size() - 1;
            } else if (kind == 3) { sf.kind = 1; sf.center[0] = specials[rng() % 12]; sf.radius_or_height = specials[rng() % 12]; }
            else { sf.kind = 2; sf.radius_or_height = specials[rng() % 12]; }
            sc.surfaces.push_back(sf);
            sc.objects.push_back(HostObject{(uint32_t)s, 0});
        }
        float eye[3] = {0, 1, 5}, ce[3] = {0, 0, 0}, up[3] = {0, 1, 0};
        memcpy(sc.camera.eye, eye, 12); camera_look_at(eye, ce, up, sc.camera.direction, sc.camera.up); sc.camera.fov = 0.5f; sc.camera.has_dof = 0;
        FlatScene flat; std::string err;
        if (!flatten_scene(sc, flat, err)) { printf("rejected: %s\n", err.c_str()); continue; }
        // structural checks: every triangle appears once, node children in range, depth within the stack
        if (flat.bvh_depth > 32) { printf("DEPTH %u\n", flat.bvh_depth); return 1; }
        size_t n_nodes = flat.nodes.size() / NODE_QUADS;
        std::vector<int> seen(flat.n_tris, 0);
        std::vector<size_t> stack{0};
        size_t visited = 0;
        while (!stack.empty()) {
            size_t k = stack.back(); stack.pop_back();
            if (++visited > n_nodes + 1) { printf("CYCLE\n"); return 1; }
            for (int c = 0; c < 2; ++c) {
                uint32_t u; float f = c == 0 ? flat.nodes[k * NODE_QUADS + 1].z : flat.nodes[k * NODE_QUADS + 1].w; memcpy(&u, &f, 4);
                int32_t code = (int32_t)u;
                if (code >= 0) { if ((size_t)code >= n_nodes) { printf("CHILD OUT OF RANGE\n"); return 1; } stack.push_back((size_t)code); }
                else { uint32_t v = ~u; uint32_t first = v >> 3, cnt = v & 7; if (first + cnt > flat.n_tris) { printf("LEAF OUT OF RANGE\n"); return 1; } for (uint32_t i = 0; i < cnt; ++i) seen[first + i]++; }
            }
        }
        for (uint32_t i = 0; i < flat.n_tris; ++i) if (seen[i] != 1) { printf("TRIANGLE %u seen %d times (n=%u)\n", i, seen[i], flat.n_tris); return 1; }
        // the reference's scene-level tree (layout.h: SceneTreeNode): pre-order, every surface in exactly one leaf,
        // skip links forward and inside the array, parents before their children
        {
            const size_t n_surfaces = flat.surface_node.size(), n_st = flat.scene_tree.size();
            if (n_st != (n_surfaces >= 2 ? 2 * n_surfaces - 1 : 0)) { printf("SCENE TREE SIZE %zu for %zu surfaces\n", n_st, n_surfaces); return 1; }
            std::vector<int> leaf_seen(n_surfaces, 0);
            for (size_t k = 0; k < n_st; ++k) {
                const SceneTreeNode& nd = flat.scene_tree[k];
                if (nd.parent >= (int32_t)k || (k == 0) != (nd.parent < 0)) { printf("SCENE TREE PARENT\n"); return 1; }
                if (nd.a < 0) {
                    const uint32_t sfc = (uint32_t)~nd.a;
                    if (sfc >= n_surfaces || flat.surface_node[sfc] != k) { printf("SCENE TREE LEAF\n"); return 1; }
                    leaf_seen[sfc]++;
                } else if ((size_t)nd.a <= k + 2 || (size_t)nd.a > n_st) { printf("SCENE TREE SKIP\n"); return 1; }
            }
            if (n_st) for (size_t sfc = 0; sfc < n_surfaces; ++sfc) if (leaf_seen[sfc] != 1) { printf("SCENE TREE SURFACE %zu\n", sfc); return 1; }
        }
    }
    printf("ok\n");
}
