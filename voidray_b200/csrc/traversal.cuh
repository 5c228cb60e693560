// traversal.cuh — closest-hit traversal of kernels.cu: Möller–Trumbore, analytic surfaces, the quantised-node slab
// tests and the short stack. Device functions only, textually part of kernels.cu (the one place that includes it);
// kept in a header of its own so that tests/c/trav_host.cpp can compile this very source for the CPU
// (-DVR_HOST_SHIM: tests/c/host_shim.h stands in for the CUDA intrinsics) and check its hits bit for bit without a
// GPU.
#pragma once
#include "device_math.cuh"
#include "layout.h"

namespace vr {

// Per-thread traversal stack: 32 entries in shared memory; the builder caps the BVH depth at STACK_DEPTH.
static constexpr int STACK_DEPTH = 32;
static constexpr int SMEM_STACK = STACK_DEPTH;
static constexpr float T_MIN = 0.00001f;  // core/scene.rs:183
static constexpr int SENTINEL = 0x7FFFFFFF;

// ------------------------------------------------------------------------------------------------
// Closest hit: core/scene.rs:182-185 semantics — the smallest t > 1e-5 over every primitive whose
// Möller–Trumbore / analytic test accepts the ray; equal t resolved by the reference's in-order rank
// (largest wins). Box tests only cull: they are padded so that they never reject a primitive the
// exact test would accept.
// ------------------------------------------------------------------------------------------------
struct HitResult {
    float t;
    int prim;  // GPU primitive: [0, n_tris) triangle, n_tris + k analytic k, -1 miss
    float u, v;
};

// ------------------------------------------------------------------------------------------------
// Scene-level culling of the reference (core/scene.rs:182-185 -> core/bvh.rs:132-160 over the surfaces): a surface is
// visited iff every Split above it passes AABB::hit. Those boxes are unions of un-expanded Mesh::bounds()
// (scene.rs:73-80, mesh.rs:92-104), so unlike the boxes of this library's own BVH they are part of the result: a
// Split that is flat on an axis (coplanar quads as separate surfaces) rejects every ray with a component along it.
// Restated exactly — same operations and order, 1 / dir, t_max <= t_min — and applied as a filter on the winning
// candidates: a ray carries one visibility bit per surface (up to SCENE_MASK_SURFACES surfaces, computed when the ray
// starts and parked in the row above its shared-memory stack), beyond that a candidate walks its own ancestor chain.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool scene_box_hit(const SceneTreeNode* __restrict__ node, f3 o, f3 inv) {  // util/aabb.rs:86-148
    const float8 n = ldg8((const float4*)node);
    const float lox = n.lo.x, loy = n.lo.y, loz = n.lo.z, hix = n.lo.w, hiy = n.hi.x, hiz = n.hi.y;
    if (lox == INFINITY && loy == INFINITY && loz == INFINITY && hix == -INFINITY && hiy == -INFINITY && hiz == -INFINITY)
        return true;  // *self == Self::default()
    float t_min = 0.00001f, t_max = INFINITY;
    {
        float t0 = (lox - o.x) * inv.x, t1 = (hix - o.x) * inv.x;
        if (inv.x < 0.0f) { const float s = t0; t0 = t1; t1 = s; }
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (loy - o.y) * inv.y, t1 = (hiy - o.y) * inv.y;
        if (inv.y < 0.0f) { const float s = t0; t0 = t1; t1 = s; }
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    {
        float t0 = (loz - o.z) * inv.z, t1 = (hiz - o.z) * inv.z;
        if (inv.z < 0.0f) { const float s = t0; t0 = t1; t1 = s; }
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    return true;
}
// One bit per surface the reference's scene tree visits for this ray (n_surfaces <= SCENE_MASK_SURFACES): pre-order
// walk, a rejected Split skips its subtree.
__device__ __forceinline__ uint32_t scene_visible_mask(const SceneTreeNode* __restrict__ tree, uint32_t n_nodes, f3 o, f3 d) {
    const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    uint32_t vis = 0u;
    for (uint32_t i = 0; i < n_nodes;) {
        const int a = tree[i].a;
        if (a < 0) {
            vis |= 1u << (~a);
            ++i;
        } else if (scene_box_hit(tree + i, o, inv)) {
            ++i;
        } else {
            i = (uint32_t)a;
        }
    }
    return vis;
}
// The same for one surface, from its leaf up to the root (scenes with more surfaces than mask bits).
__device__ __forceinline__ bool scene_surface_visible(const SceneTreeNode* __restrict__ tree, uint32_t leaf, f3 o, f3 d) {
    const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    for (int i = tree[leaf].parent; i >= 0; i = tree[i].parent)
        if (!scene_box_hit(tree + i, o, inv)) return false;
    return true;
}
// Is a candidate on `surface` one the reference would have tested? vis_row = the ray's parked visibility word.
__device__ __forceinline__ bool candidate_visible(const DeviceScene& sc, const int* vis_row, uint32_t surface, f3 o, f3 d) {
    if (sc.n_scene_nodes == 0u) return true;
    if (sc.n_surfaces <= SCENE_MASK_SURFACES) return (((uint32_t)*vis_row) >> surface) & 1u;
    return scene_surface_visible(sc.scene_tree, sc.surface_node[surface], o, d);
}

__device__ __forceinline__ void intersect_triangle(const DeviceScene& sc, const int* vis_row,
                                                   const float4* __restrict__ tri_isect, int tri, f3 o, f3 d,
                                                   HitResult& best, uint32_t& best_rank) {
    const float8 r0 = ldg8_once(tri_isect + TRI_ISECT_QUADS * tri);      // v0 | e1
    const float8 r1 = ldg8_once(tri_isect + TRI_ISECT_QUADS * tri + 2);  // e2 | -
    const float4 q0 = r0.lo;
    const f3 v0 = xyz(r0.lo), e1 = xyz(r0.hi), e2 = xyz(r1.lo);
    // core/mesh.rs:153-175, same operation order
    const f3 h = cross(d, e2);
    const float a = dot(e1, h);
    if (a > -T_MIN && a < T_MIN) return;
    const float f = 1.0f / a;
    const f3 s = o - v0;
    const float u = f * dot(s, h);
    if (u < 0.0f || u > 1.0f) return;
    const f3 q = cross(s, e1);
    const float v = f * dot(d, q);
    if (v < 0.0f || u + v > 1.0f) return;
    const float t = f * dot(e2, q);
    if (t > T_MIN) {
        const uint32_t rank = __float_as_uint(q0.w);
        if (t < best.t || (t == best.t && rank > best_rank)) {
            // would win: was its surface reached by the reference's scene tree?
            if (sc.n_scene_nodes != 0u) {
                const uint32_t surface = __float_as_uint(r1.hi.w);  // (the record bypassed the L1: no cheap re-read)
                if (!candidate_visible(sc, vis_row, surface, o, d)) return;
            }
            best.t = t;
            best.prim = tri;
            best.u = u;
            best.v = v;
            best_rank = rank;
        }
    }
}

__device__ __forceinline__ void intersect_analytic(const DeviceScene& sc, const int* vis_row, const AnalyticRec& a,
                                                   int prim, f3 o, f3 d, HitResult& best, uint32_t& best_rank) {
    float t;
    if (a.kind == 0) {  // Sphere::hit, voidray_common/src/surfaces.rs:46-80 with (t_min, t_max) = (1e-5, inf)
        const f3 oc = o - mk3(a.cx, a.cy, a.cz);
        const float aa = magnitude2(d);
        const float half_b = dot(oc, d);
        const float c = magnitude2(oc) - a.radius * a.radius;
        const float disc = half_b * half_b - aa * c;
        if (disc < 0.0f) return;
        const float sqrtd = sqrtf(disc);
        float root = (-half_b - sqrtd) / aa;
        if (root < T_MIN || INFINITY < root) {
            root = (-half_b + sqrtd) / aa;
            if (root < T_MIN || INFINITY < root) return;
        }
        t = root;
    } else {  // GroundPlane::hit, surfaces.rs:87-105
        t = (a.radius - o.y) / d.y;
        if (t <= T_MIN || t >= INFINITY) return;
    }
    if (t < best.t || (t == best.t && a.rank > best_rank) || best.prim < 0) {
        // (best.prim < 0 covers a NaN-free first hit at t == inf, which the tests above exclude anyway)
        if (!candidate_visible(sc, vis_row, a.surface, o, d)) return;
        best.t = t;
        best.prim = prim;
        best.u = 0.0f;
        best.v = 0.0f;
        best_rank = a.rank;
    }
}

// Per-lane traversal state. The SMEM_STACK-entry stack lives in shared memory (one column per thread,
// stride = blockDim.x, conflict-free); the builder caps the BVH depth so it cannot overflow.
struct Traversal {
    f3 o, d;
    // Slab tests on the quantised nodes (layout.h), never feeding a reported value:
    //   t = fma(f, a, b),  f = 1 + q / 32768 (decoded by one PRMT),  a = extent * id,  b = (grid_min - o) * id - a.
    // The near plane (q_lo if id >= 0, else q_hi) uses bn = b - err, the far plane bf = b + err, where err bounds the
    // rounding of b per axis (it only ever touches that axis' distances, so a ray with a tiny direction component
    // keeps culling on the other two axes — unlike a per-ray slack). No min / max per axis is needed.
    float ax, ay, az, bnx, bny, bnz, bfx, bfy, bfz;
    uint32_t selx, sely, selz;  // PRMT selector of the near plane's half-word; far = sel ^ 0x0220
    HitResult best;
    uint32_t best_rank;
    int cur, sp;
};
__device__ __forceinline__ void trav_axis(float o, float d, float gmin, float extent, float& a, float& bn, float& bf,
                                          uint32_t& sel) {
    const float tiny = 1e-20f;
    const float id = 1.0f / (fabsf(d) > tiny ? d : copysignf(tiny, d));
    a = extent * id;
    const float g = (gmin - o) * id;
    const float b = g - a;
    // 2.4e-7 (|g| + |a|) bounds the rounding of b = g - a and of g itself; the plane's own FFMA t = f * a + b (f in
    // [1, 2)) rounds by at most 2^-24 (2 |a| + |b|) <= 6e-8 (|g| + 3 |a|): both are folded into the addends, so the
    // node step compares near <= far directly, with no widening multiply
    const float err = 4.5e-7f * (fabsf(g) + fabsf(a)) + 1e-30f;
    bn = b - err;
    bf = b + err;
    sel = id >= 0.0f ? 0x7104u : 0x7324u;  // bytes (0x00, q.b0, q.b1, 0x3F) of the low / high half-word
}

// The row above a thread's shared-memory stack holds the ray's scene-level visibility word (candidate_visible).
__device__ __forceinline__ const int* trav_vis_row(const int* sstack, int sstride) { return sstack + SMEM_STACK * sstride; }

__device__ __forceinline__ void trav_begin(Traversal& tr, const DeviceScene& sc, f3 o, f3 d, int* sstack, int sstride) {
    tr.o = o;
    tr.d = d;
    if (sc.n_scene_nodes != 0u && sc.n_surfaces <= SCENE_MASK_SURFACES)
        sstack[SMEM_STACK * sstride] = (int)scene_visible_mask(sc.scene_tree, sc.n_scene_nodes, o, d);
    trav_axis(o.x, d.x, sc.grid_min[0], sc.grid_extent[0], tr.ax, tr.bnx, tr.bfx, tr.selx);
    trav_axis(o.y, d.y, sc.grid_min[1], sc.grid_extent[1], tr.ay, tr.bny, tr.bfy, tr.sely);
    trav_axis(o.z, d.z, sc.grid_min[2], sc.grid_extent[2], tr.az, tr.bnz, tr.bfz, tr.selz);
    tr.best.t = INFINITY;
    tr.best.prim = -1;
    tr.best.u = tr.best.v = 0.0f;
    tr.best_rank = 0;
    tr.sp = 0;
    tr.cur = sc.n_tris > 0 ? 0 : SENTINEL;
}

// The stack lives entirely in shared memory (one column per thread, stride = blockDim.x: conflict-free).
// Push and pop are written so that they compile to predicated STS / LDS instead of branches.
__device__ __forceinline__ int trav_pop(Traversal& tr, const int* sstack, int sstride) {
    const bool empty = tr.sp == 0;
    tr.sp -= empty ? 0 : 1;
    const int v = sstack[tr.sp * sstride];
    return empty ? SENTINEL : v;
}

__device__ __forceinline__ bool is_inner(int cur) { return (unsigned)cur < (unsigned)SENTINEL; }  // leaf codes are negative

// Plane distance from a packed (q_lo | q_hi << 16) word: PRMT builds f = 1 + q / 32768, one FFMA maps it to t.
__device__ __forceinline__ float plane_t(uint32_t pair, uint32_t sel, float a, float b) {
    return __fmaf_rn(__uint_as_float(__byte_perm(pair, 0x3F000000u, sel)), a, b);
}

// One inner node: two slab tests from a single 32-byte record, near child first, far child pushed.
__device__ __forceinline__ void trav_node(Traversal& tr, const float4* __restrict__ nodes, int* sstack, int sstride) {
    const float8 n = ldg8_keep((const float4*)((const char*)nodes + (size_t)(uint32_t)tr.cur * 32u));  // one IMAD.WIDE
    const uint32_t w0 = __float_as_uint(n.lo.x), w1 = __float_as_uint(n.lo.y), w2 = __float_as_uint(n.lo.z),
                   w3 = __float_as_uint(n.lo.w), w4 = __float_as_uint(n.hi.x), w5 = __float_as_uint(n.hi.y);
    const uint32_t fx = tr.selx ^ 0x0220u, fy = tr.sely ^ 0x0220u, fz = tr.selz ^ 0x0220u;
    const float an = fmaxf(fmaxf(plane_t(w0, tr.selx, tr.ax, tr.bnx), plane_t(w1, tr.sely, tr.ay, tr.bny)),
                           fmaxf(plane_t(w2, tr.selz, tr.az, tr.bnz), 0.0f));
    const float af = fminf(fminf(plane_t(w0, fx, tr.ax, tr.bfx), plane_t(w1, fy, tr.ay, tr.bfy)),
                           fminf(plane_t(w2, fz, tr.az, tr.bfz), tr.best.t));
    const float bn = fmaxf(fmaxf(plane_t(w3, tr.selx, tr.ax, tr.bnx), plane_t(w4, tr.sely, tr.ay, tr.bny)),
                           fmaxf(plane_t(w5, tr.selz, tr.az, tr.bnz), 0.0f));
    const float bf = fminf(fminf(plane_t(w3, fx, tr.ax, tr.bfx), plane_t(w4, fy, tr.ay, tr.bfy)),
                           fminf(plane_t(w5, fz, tr.az, tr.bfz), tr.best.t));
    // conservative: the boxes carry a guard cell, and bn / bf carry the rounding of the addend and of the FFMA
    // itself (trav_axis; Ize, "Robust BVH ray traversal", 2013)
    const bool hit_a = an <= af;
    const bool hit_b = bn <= bf;
    const int ca = __float_as_int(n.hi.z), cb = __float_as_int(n.hi.w);
    // branch-free child selection: the divergent if/else ladder ran at 2-3 lanes per instruction
    const bool b_first = hit_b && (!hit_a || bn < an);
    const int near_c = b_first ? cb : ca;
    const int far_c = b_first ? ca : cb;
    const bool both = hit_a && hit_b, any = hit_a || hit_b;
    if (both) sstack[tr.sp * sstride] = far_c;
    tr.sp += both ? 1 : 0;
    int next = near_c;
    if (!any) next = trav_pop(tr, sstack, sstride);
    tr.cur = next;
}


// One triangle of the current leaf; the leaf code counts down so that lanes with short leaves do not idle
// through a neighbour's longer one.
__device__ __forceinline__ void trav_leaf_step(Traversal& tr, const DeviceScene& sc, const float4* __restrict__ tri_isect,
                                               int* sstack, int sstride) {
    const int code = ~tr.cur;
    const int first = code >> 3, count = code & 7;
    if (count > 0) intersect_triangle(sc, trav_vis_row(sstack, sstride), tri_isect, first, tr.o, tr.d, tr.best, tr.best_rank);
    int next = ~(((first + 1) << 3) | (count - 1));
    if (count <= 1) next = trav_pop(tr, sstack, sstride);
    tr.cur = next;
}

__device__ __forceinline__ HitResult trav_finish(Traversal& tr, const DeviceScene& sc, const int* sstack, int sstride) {
    for (uint32_t k = 0; k < sc.n_analytics; ++k)
        intersect_analytic(sc, trav_vis_row(sstack, sstride), sc.analytics[k], (int)(sc.n_tris + k), tr.o, tr.d, tr.best,
                           tr.best_rank);
    return tr.best;
}

// One ray, start to finish (gate kernels). sstack: SMEM_STACK + 1 rows of sstride ints (stack + visibility word).
__device__ __forceinline__ HitResult closest_hit(const DeviceScene& sc, f3 o, f3 d, int* sstack, int sstride) {
    Traversal tr;
    trav_begin(tr, sc, o, d, sstack, sstride);
    const float4* __restrict__ nodes = (const float4*)sc.nodes;
    const float4* __restrict__ tri_isect = (const float4*)sc.tri_isect;
    while (tr.cur != SENTINEL) {
        if (is_inner(tr.cur)) trav_node(tr, nodes, sstack, sstride);
        else trav_leaf_step(tr, sc, tri_isect, sstack, sstride);
    }
    return trav_finish(tr, sc, sstack, sstride);
}

}  // namespace vr
