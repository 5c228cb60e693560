// traversal.cuh — closest-hit traversal of kernels.cu: Möller–Trumbore, analytic surfaces, the quantised-node slab
// tests and the short stack. Device functions only, textually part of kernels.cu (the one place that includes it);
// kept in a header of its own so that tests/c/trav_host.cpp can compile this very source for the CPU
// (-DVR_HOST_SHIM: tests/c/host_shim.h stands in for the CUDA intrinsics) and check its hits bit for bit without a
// GPU — in the shipped layout and in every experiment variant (-DVR_BVH4, -DVR_TRI48, -DVR_SMEM_STACK).
#pragma once
#include "device_math.cuh"
#include "layout.h"

namespace vr {

// Per-thread traversal stack: VR_SMEM_STACK entries in shared memory; the builder caps the BVH depth at STACK_DEPTH.
// Experiment -DVR_SMEM_STACK=16: half the shared memory per block (more of the SM's 256 KB left to the L1; the L1
// model of scripts/bvh_stats.cpp gives 64 -> 128 KB about 8 points of sector hit rate on bounce rays), entries past
// the shared part go to a per-thread local array (the deepest stack seen on the BASELINE scenes is 13).
#ifndef VR_SMEM_STACK
#define VR_SMEM_STACK 32
#endif
// Experiment -DVR_BVH4 (layout.h): 4-wide nodes park up to three children per step; the host collapse narrows nodes
// where needed so that no path can park more than WIDE_STACK_LIMIT = 32 entries, the same stack as the BVH2.
static constexpr int STACK_DEPTH = 32;
#ifdef VR_BVH4
static_assert(WIDE_STACK_LIMIT == STACK_DEPTH, "the collapse must bound the stack the kernel has");
#endif
#if VR_SMEM_STACK < 32
#define VR_HAS_SPILL 1
#endif
static constexpr int SMEM_STACK = VR_SMEM_STACK;
static_assert(SMEM_STACK >= 1 && SMEM_STACK <= STACK_DEPTH, "VR_SMEM_STACK");
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

__device__ __forceinline__ void intersect_triangle(const float4* __restrict__ tri_isect, int tri, f3 o, f3 d,
                                                   HitResult& best, uint32_t& best_rank) {
#ifdef VR_TRI48
    const float4 q0 = ldg4(tri_isect + TRI_ISECT_QUADS * tri);  // 48-byte records: only 16-byte aligned
    const f3 v0 = xyz(q0), e1 = xyz(ldg4(tri_isect + TRI_ISECT_QUADS * tri + 1)),
             e2 = xyz(ldg4(tri_isect + TRI_ISECT_QUADS * tri + 2));
#else
    const float8 r0 = ldg8(tri_isect + TRI_ISECT_QUADS * tri);      // v0 | e1
    const float8 r1 = ldg8(tri_isect + TRI_ISECT_QUADS * tri + 2);  // e2 | -
    const float4 q0 = r0.lo;
    const f3 v0 = xyz(r0.lo), e1 = xyz(r0.hi), e2 = xyz(r1.lo);
#endif
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
            best.t = t;
            best.prim = tri;
            best.u = u;
            best.v = v;
            best_rank = rank;
        }
    }
}

__device__ __forceinline__ void intersect_analytic(const AnalyticRec& a, int prim, f3 o, f3 d, HitResult& best,
                                                   uint32_t& best_rank) {
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
// VR_SMEM_STACK < 32: entries SMEM_STACK.. of the stack live in a per-thread local array that is passed alongside
// the shared part (kept out of Traversal: a dynamically indexed member would drag the whole struct into local memory)
#ifdef VR_HAS_SPILL
#define VR_SPILL_PARAM , int* __restrict__ spill
#define VR_SPILL_ARG , spill
#define VR_SPILL_DECL int spill[STACK_DEPTH - SMEM_STACK];
#else
#define VR_SPILL_PARAM
#define VR_SPILL_ARG
#define VR_SPILL_DECL
#endif
// Experiment -DVR_TRACE_SPEC -DVR_SPEC_ARRIVAL (kernels.cu): the node step itself parks a leaf it arrives at when the
// caller passes a free parking slot; the gate kernels pass none and the parameter folds away.
#if defined(VR_TRACE_SPEC) && defined(VR_SPEC_ARRIVAL)
#define VR_PARK_PARAM , int* __restrict__ park
#define VR_PARK_NONE , nullptr
#else
#define VR_PARK_PARAM
#define VR_PARK_NONE
#endif

__device__ __forceinline__ void trav_axis(float o, float d, float gmin, float extent, float& a, float& bn, float& bf,
                                          uint32_t& sel) {
    const float tiny = 1e-20f;
    const float id = 1.0f / (fabsf(d) > tiny ? d : copysignf(tiny, d));
    a = extent * id;
    const float g = (gmin - o) * id;
    const float b = g - a;
    const float err = 2.4e-7f * (fabsf(g) + fabsf(a)) + 1e-30f;
    bn = b - err;
    bf = b + err;
    sel = id >= 0.0f ? 0x7104u : 0x7324u;  // bytes (0x00, q.b0, q.b1, 0x3F) of the low / high half-word
}

__device__ __forceinline__ void trav_begin(Traversal& tr, const DeviceScene& sc, f3 o, f3 d) {
    tr.o = o;
    tr.d = d;
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
__device__ __forceinline__ int trav_pop(Traversal& tr, const int* sstack, int sstride VR_SPILL_PARAM) {
    const bool empty = tr.sp == 0;
    tr.sp -= empty ? 0 : 1;
#ifdef VR_HAS_SPILL
    const int v = tr.sp < SMEM_STACK ? sstack[tr.sp * sstride] : spill[tr.sp - SMEM_STACK];
#else
    const int v = sstack[tr.sp * sstride];
#endif
    return empty ? SENTINEL : v;
}

__device__ __forceinline__ bool is_inner(int cur) { return (unsigned)cur < (unsigned)SENTINEL; }  // leaf codes are negative

// Plane distance from a packed (q_lo | q_hi << 16) word: PRMT builds f = 1 + q / 32768, one FFMA maps it to t.
__device__ __forceinline__ float plane_t(uint32_t pair, uint32_t sel, float a, float b) {
    return __fmaf_rn(__uint_as_float(__byte_perm(pair, 0x3F000000u, sel)), a, b);
}

#ifdef VR_BVH4
// One slab test on the three packed words of a child; same arithmetic as the BVH2 step below.
__device__ __forceinline__ float wide_child(const Traversal& tr, uint32_t wx, uint32_t wy, uint32_t wz) {
    const uint32_t fx = tr.selx ^ 0x0220u, fy = tr.sely ^ 0x0220u, fz = tr.selz ^ 0x0220u;
    const float tn = fmaxf(fmaxf(plane_t(wx, tr.selx, tr.ax, tr.bnx), plane_t(wy, tr.sely, tr.ay, tr.bny)),
                           fmaxf(plane_t(wz, tr.selz, tr.az, tr.bnz), 0.0f));
    const float tf = fminf(fminf(plane_t(wx, fx, tr.ax, tr.bfx), plane_t(wy, fy, tr.ay, tr.bfy)),
                           fminf(plane_t(wz, fz, tr.az, tr.bfz), tr.best.t));
    return tn <= tf * 1.0000005f ? tn : INFINITY;  // the sort key: entry distance, +inf for a miss
}
__device__ __forceinline__ void wide_cswap(float& ka, int& ca, float& kb, int& cb) {
    const bool s = kb < ka;
    const float k0 = s ? kb : ka, k1 = s ? ka : kb;
    const int c0 = s ? cb : ca, c1 = s ? ca : cb;
    ka = k0;
    kb = k1;
    ca = c0;
    cb = c1;
}
__device__ __forceinline__ void wide_push(Traversal& tr, int* sstack, int sstride VR_SPILL_PARAM, bool pred, int v) {
#ifdef VR_HAS_SPILL
    if (pred) {
        if (tr.sp < SMEM_STACK) sstack[tr.sp * sstride] = v;
        else spill[tr.sp - SMEM_STACK] = v;
    }
#else
    if (pred) sstack[tr.sp * sstride] = v;
#endif
    tr.sp += pred ? 1 : 0;
}
// One 4-wide node: two 256-bit loads, four slab tests, children that are hit sorted by entry distance (a 5-comparator
// network, branch-free); the nearest is next, the others are parked farthest first. scripts/bvh_stats.cpp walks the
// same node array with the same arithmetic on the CPU: half the node fetches of the BVH2 for the same triangle tests.
__device__ __forceinline__ void trav_node(Traversal& tr, const float4* __restrict__ nodes, int* sstack, int sstride VR_SPILL_PARAM VR_PARK_PARAM) {
#if defined(VR_TRACE_SPEC) && defined(VR_SPEC_ARRIVAL)
    (void)park;  // arrival parking is only built for the binary step
#endif
    const float8 p0 = ldg8(nodes + WIDE_NODE_QUADS * tr.cur);
    const float8 p1 = ldg8(nodes + WIDE_NODE_QUADS * tr.cur + 2);
    float k0 = wide_child(tr, __float_as_uint(p0.lo.x), __float_as_uint(p0.lo.y), __float_as_uint(p0.lo.z));
    float k1 = wide_child(tr, __float_as_uint(p0.lo.w), __float_as_uint(p0.hi.x), __float_as_uint(p0.hi.y));
    float k2 = wide_child(tr, __float_as_uint(p1.lo.x), __float_as_uint(p1.lo.y), __float_as_uint(p1.lo.z));
    float k3 = wide_child(tr, __float_as_uint(p1.lo.w), __float_as_uint(p1.hi.x), __float_as_uint(p1.hi.y));
    int c0 = __float_as_int(p0.hi.z), c1 = __float_as_int(p0.hi.w), c2 = __float_as_int(p1.hi.z), c3 = __float_as_int(p1.hi.w);
    const int hits = (k0 < INFINITY ? 1 : 0) + (k1 < INFINITY ? 1 : 0) + (k2 < INFINITY ? 1 : 0) + (k3 < INFINITY ? 1 : 0);
#ifdef VR_BVH4_NOSORT
    // Experiment: only the nearest child is found (three compare-selects instead of five), the other hits are parked
    // in slot order. The CPU walk gives +1 % node fetches and triangle tests for it (BVH_STATS_NOSORT=1).
    wide_cswap(k0, c0, k1, c1);
    wide_cswap(k0, c0, k2, c2);
    wide_cswap(k0, c0, k3, c3);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, k3 < INFINITY, c3);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, k2 < INFINITY, c2);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, k1 < INFINITY, c1);
#else
    wide_cswap(k0, c0, k1, c1);
    wide_cswap(k2, c2, k3, c3);
    wide_cswap(k0, c0, k2, c2);
    wide_cswap(k1, c1, k3, c3);
    wide_cswap(k1, c1, k2, c2);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, hits > 3, c3);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, hits > 2, c2);
    wide_push(tr, sstack, sstride VR_SPILL_ARG, hits > 1, c1);
#endif
    int next = c0;
    if (hits == 0) next = trav_pop(tr, sstack, sstride VR_SPILL_ARG);
    tr.cur = next;
}
#else
// One inner node: two slab tests from a single 32-byte record, near child first, far child pushed.
__device__ __forceinline__ void trav_node(Traversal& tr, const float4* __restrict__ nodes, int* sstack, int sstride VR_SPILL_PARAM VR_PARK_PARAM) {
    const float8 n = ldg8(nodes + 2 * tr.cur);
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
    // conservative: the boxes carry a guard cell, bn / bf carry the addend's rounding, and the exit is widened
    // by a few ulps for the FFMA's own rounding (Ize, "Robust BVH ray traversal", 2013)
    const bool hit_a = an <= af * 1.0000005f;
    const bool hit_b = bn <= bf * 1.0000005f;
    const int ca = __float_as_int(n.hi.z), cb = __float_as_int(n.hi.w);
    // branch-free child selection: the divergent if/else ladder ran at 2-3 lanes per instruction
    const bool b_first = hit_b && (!hit_a || bn < an);
    const int near_c = b_first ? cb : ca;
    const int far_c = b_first ? ca : cb;
    const bool both = hit_a && hit_b, any = hit_a || hit_b;
#ifdef VR_HAS_SPILL
    if (both) {
        if (tr.sp < SMEM_STACK) sstack[tr.sp * sstride] = far_c;
        else spill[tr.sp - SMEM_STACK] = far_c;
    }
#else
    if (both) sstack[tr.sp * sstride] = far_c;
#endif
    tr.sp += both ? 1 : 0;
    int next = near_c;
#if defined(VR_TRACE_SPEC) && defined(VR_SPEC_ARRIVAL)
    // arrival at a leaf with a free parking slot and something left on the stack: park it and walk on (shares the pop)
    const bool park_it = park != nullptr && any && near_c < 0 && *park == SENTINEL && tr.sp > 0;
    if (park_it) *park = near_c;
    if (!any || park_it) next = trav_pop(tr, sstack, sstride VR_SPILL_ARG);
#else
    if (!any) next = trav_pop(tr, sstack, sstride VR_SPILL_ARG);
#endif
    tr.cur = next;
}

#endif  // VR_BVH4

// One triangle of the current leaf; the leaf code counts down so that lanes with short leaves do not idle
// through a neighbour's longer one.
__device__ __forceinline__ void trav_leaf_step(Traversal& tr, const float4* __restrict__ tri_isect, int* sstack,
                                               int sstride VR_SPILL_PARAM) {
#ifdef VR_LEAF_COMPACT
    // Experiment (register pressure, used with -DVR_TRACE_CHUNK): nothing but tr.cur lives across the triangle test;
    // (first + 1) << 3 | (count - 1) is the packed code plus 7.
    if (((~tr.cur) & 7) > 0) intersect_triangle(tri_isect, (~tr.cur) >> 3, tr.o, tr.d, tr.best, tr.best_rank);
    const int code = ~tr.cur;
    int next = ~(code + 7);
    if ((code & 7) <= 1) next = trav_pop(tr, sstack, sstride VR_SPILL_ARG);
    tr.cur = next;
#else
    const int code = ~tr.cur;
    const int first = code >> 3, count = code & 7;
    if (count > 0) intersect_triangle(tri_isect, first, tr.o, tr.d, tr.best, tr.best_rank);
    int next = ~(((first + 1) << 3) | (count - 1));
    if (count <= 1) next = trav_pop(tr, sstack, sstride VR_SPILL_ARG);
    tr.cur = next;
#endif
}

__device__ __forceinline__ HitResult trav_finish(Traversal& tr, const DeviceScene& sc) {
    for (uint32_t k = 0; k < sc.n_analytics; ++k)
        intersect_analytic(sc.analytics[k], (int)(sc.n_tris + k), tr.o, tr.d, tr.best, tr.best_rank);
    return tr.best;
}

// One ray, start to finish (gate kernels).
__device__ __forceinline__ HitResult closest_hit(const DeviceScene& sc, f3 o, f3 d, int* sstack, int sstride) {
    Traversal tr;
    VR_SPILL_DECL
    trav_begin(tr, sc, o, d);
    const float4* __restrict__ nodes = (const float4*)sc.nodes;
    const float4* __restrict__ tri_isect = (const float4*)sc.tri_isect;
    while (tr.cur != SENTINEL) {
        if (is_inner(tr.cur)) trav_node(tr, nodes, sstack, sstride VR_SPILL_ARG VR_PARK_NONE);
        else trav_leaf_step(tr, tri_isect, sstack, sstride VR_SPILL_ARG);
    }
    return trav_finish(tr, sc);
}

}  // namespace vr
