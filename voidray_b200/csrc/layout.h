// layout.h — packed device layouts of the flattened scene and of the wavefront state.
// Everything the kernels fetch is a 16-byte float4 (one LDG.128 per lane); see DESIGN.md §3.
#pragma once
#include <stdint.h>

namespace vr {

// BVH2 node, 32 B = one 256-bit load. A node stores the boxes of BOTH children, quantised to a 15-bit grid
// over the scene bounds (cell = extent / 32768; lo rounded down, hi rounded up, plus one guard cell):
//   w0..w2 = child 0 (x, y, z), w3..w5 = child 1 (x, y, z): each word = (0x8000 | q_lo) | (0x8000 | q_hi) << 16
//   w6, w7 = child codes: >= 0 inner node index; < 0 leaf, ~c = (first_tri << 3) | count
// The stored 16-bit value already carries the float's implicit-one position: PRMT drops its two bytes into bytes
// 1..2 of 0x3F000000, giving f = 1 + q / 32768 exactly, and the slab distance is one FFMA per plane,
//   t = f * (extent * id) + ((grid_min - o) * id - extent * id),
// with the near / far plane picked by the ray's sign (PRMT selector) instead of min / max.
static const int NODE_QUADS = 2;
static const int DEVICE_NODE_QUADS = NODE_QUADS;
static const int LEAF_MAX_TRIS = 4;

// Intersection record, 64 B = 4 x float4 = two 256-bit loads (pre-subtracted edges: e1 = v1 - v0,
// e2 = v2 - v0 are the same f32 subtractions core/mesh.rs:150-151 performs per test, done once on the host).
//   q0 = (v0.xyz, tie rank as uint bits)
//   q1 = (e1.xyz, 0)
//   q2 = (e2.xyz, 0)
//   q3 = (0, 0, 0, surface handle as uint bits): padding to the 32-byte alignment the 256-bit loads need; the
//        surface handle is read only when a candidate would win and the scene has a scene-level tree (below)
static const int TRI_ISECT_QUADS = 4;

// Shading record, 80 B = 5 x float4, fetched once per closest hit.
//   q0 = (n0.xyz, uv0.x)   q1 = (n1.xyz, uv0.y)   q2 = (n2.xyz, uv1.x)
//   q3 = (ng.xyz, uv1.y)   q4 = (uv2.x, uv2.y, material index bits, 0)
static const int TRI_SHADE_QUADS = 5;

struct MaterialRec {  // 48 B
    int32_t kind;
    float color[3];
    float param;
    int32_t albedo_tex;
    int32_t normal_tex;
    int32_t transparent;
    float index, roughness, metallic, emittance;  // microfacet
};

struct TextureRec {  // texels are RGBA f32 (w unused; one 16-byte load per tap) or RGBA8 for 8-bit sources (pad = 1)
    const void* texels;  // float4*
    uint32_t width, height;
    int32_t sample_type;
    uint32_t pad;  // 1 = texels are RGBA8 (4 B; an 8-bit source), widened on the device with the exact v / 255
};

struct AnalyticRec {  // 32 B; tested linearly after the triangle BVH (scenes have a handful)
    int32_t kind;     // 0 sphere, 1 ground plane
    float cx, cy, cz; // sphere centre
    float radius;     // sphere radius | plane height
    uint32_t rank;    // tie rank in the reference's in-order sequence
    uint32_t material;
    uint32_t surface;
};

// Scene-level tree of the reference (core/scene.rs:163-179: BvhNode::from_list over the surfaces), kept for its
// culling semantics only: SceneAcceleration::hit (scene.rs:182-185) visits a surface iff every Split above it passes
// AABB::hit (util/aabb.rs:86-148), and those boxes are unions of un-expanded Mesh::bounds() — a Split that is flat on
// an axis rejects every ray with a component along it (t_max <= t_min). 32 B per node, pre-order:
//   Split: lo, hi = its box; a = index of the first node after its subtree (>= 2); parent = its Split or -1
//   leaf:  lo, hi unused;    a = ~surface handle (< 0);                         parent = its Split
// Scenes with fewer than two surfaces have no Split and no nodes.
struct SceneTreeNode {
    float lo[3], hi[3];
    int32_t a;
    int32_t parent;
};
static const uint32_t SCENE_MASK_SURFACES = 32;  // up to here a ray carries one visibility bit per surface

struct CameraRec {
    float origin[3];
    float direction[3];
    float right[3];
    float up[3];
    float d;
    int32_t has_dof;
    float aperture;
    float focal_length;
};

struct DeviceScene {
    const void* nodes;      // float4*
    const void* tri_isect;  // float4*
    const void* tri_shade;  // float4*
    const uint32_t* tri_surface;  // GPU triangle -> surface handle
    const uint32_t* tri_prim;     // GPU triangle -> triangle index inside its mesh
    const MaterialRec* materials;
    const TextureRec* textures;
    const AnalyticRec* analytics;
    const SceneTreeNode* scene_tree;  // reference scene-level tree (culling semantics), n_scene_nodes entries
    const uint32_t* surface_node;     // surface handle -> its leaf in scene_tree
    uint32_t n_scene_nodes;           // 0: fewer than two surfaces, nothing is culled at scene level
    uint32_t n_surfaces;
    float grid_min[3];     // node quantisation grid: plane(q) = grid_min + grid_extent * q / 32768
    float grid_extent[3];
    uint32_t n_tris;
    uint32_t n_analytics;
    int32_t has_microfacet;  // some material is a MicrofacetBSDF (selects the shading kernel variant)
    int32_t env_kind;  // 0 none, 1 uniform, 2 hdri
    float env_color[3];
    TextureRec env_tex;
    const float* env_marginal;  // integrator 1: P(row < j), H + 1 entries
    const float* env_cond;      // integrator 1: P(col < i | row j), H x (W + 1) entries
    CameraRec camera;
};

}  // namespace vr
