// scene_build.h — host-side scene description and flattening (the `build_acceleration` half of the
// boundary, core/scene.rs:163-179). Pure host C++; no CUDA types.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "layout.h"

namespace vr {

// std::vector whose resize() leaves trivially-constructible elements uninitialised: the large arrays of a commit
// are first touched by the worker threads that fill them instead of being zeroed by one thread.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U>
    struct rebind {
        typedef DefaultInitAllocator<U> other;
    };
    DefaultInitAllocator() = default;
    template <class U>
    DefaultInitAllocator(const DefaultInitAllocator<U>&) {}
    template <class U>
    void construct(U* p) {
        ::new ((void*)p) U;
    }
    template <class U, class... Args>
    void construct(U* p, Args&&... args) {
        ::new ((void*)p) U(std::forward<Args>(args)...);
    }
};
template <class T>
using RawVector = std::vector<T, DefaultInitAllocator<T>>;

struct Quad {
    float x, y, z, w;
};

struct HostMesh {
    std::vector<float> pos, uv, nrm;  // 3n, 2n, 3n
    std::vector<uint32_t> idx;        // 3 per triangle (trailing partial chunk already dropped)
    uint32_t n_vertices = 0;
};

struct HostSurface {
    int kind = 0;  // 0 mesh, 1 sphere, 2 ground plane
    uint32_t mesh = 0;
    float center[3] = {0, 0, 0};
    float radius_or_height = 0;
};

struct HostTexture {
    std::vector<float> rgb;  // 3*w*h
    std::vector<uint8_t> rgba8;  // 4*w*h when every float is exactly v / 255 (an 8-bit source), else empty
    uint32_t w = 0, h = 0;
    int32_t sample_type = 0;
};

struct HostObject {
    uint32_t surface, material;
};

struct HostCamera {
    float eye[3], direction[3], up[3];
    float fov;
    int32_t has_dof;
    float aperture;
    float focal_point[3];
};

struct HostScene {
    std::vector<HostMesh> meshes;
    std::vector<HostSurface> surfaces;
    std::vector<HostTexture> textures;
    std::vector<MaterialRec> materials;
    std::vector<HostObject> objects;
    HostCamera camera;
    int32_t env_kind = 0;
    float env_color[3] = {0, 0, 0};
    HostTexture env_image;
};

struct FlatScene {
    RawVector<Quad> nodes;      // NODE_QUADS per node; node 0 is the root
    RawVector<Quad> tri_isect;  // TRI_ISECT_QUADS per triangle, BVH leaf order
    RawVector<Quad> tri_shade;  // TRI_SHADE_QUADS per triangle
    RawVector<uint32_t> tri_surface, tri_prim;
    std::vector<AnalyticRec> analytics;
    std::vector<RawVector<uint32_t>> mesh_tie_rank;  // per surface (empty for analytic): rank inside the mesh
    std::vector<uint32_t> surface_rank_base;           // per surface: first global rank
    std::vector<SceneTreeNode> scene_tree;             // reference scene-level tree in pre-order (layout.h)
    std::vector<uint32_t> surface_node;                // per surface: its leaf in scene_tree (when the tree exists)
    CameraRec camera;
    float grid_min[3] = {0, 0, 0}, grid_extent[3] = {1, 1, 1};  // node quantisation grid
    uint32_t n_tris = 0;
    uint32_t bvh_depth = 0;
};

// In-order leaf sequence of the reference's median-split tree (core/bvh.rs:48-130) over items with
// the given boxes (6 floats each: min xyz, max xyz). order[i] = item at in-order position i.
void reference_leaf_order(const float* boxes6, size_t n, RawVector<uint32_t>& order);

// Camera::look_at, core/camera.rs:26-36
void camera_look_at(const float eye[3], const float center[3], const float up_in[3], float direction[3],
                    float up[3]);

bool flatten_scene(const HostScene& in, FlatScene& out, std::string& err);

// Fills t.rgba8 when every float of t.rgb is exactly v / 255 (an 8-bit source), else leaves it empty.
void pack_texture_rgba8(HostTexture& t);

// Wavefront OBJ with obj-rs 0.7.0 `load_obj::<TexturedVertex, u32>` semantics (core/mesh.rs:46-74).
bool load_obj_file(const char* path, HostMesh& out, std::string& err);

// integrator 1: luminance x sin(polar angle) sampling tables of a lat-long environment image.
// marginal: H + 1 entries, P(row < j); cond: H x (W + 1) entries, P(col < i | row j).
void build_env_tables(const HostTexture& env, std::vector<float>& marginal, std::vector<float>& cond);

}  // namespace vr
