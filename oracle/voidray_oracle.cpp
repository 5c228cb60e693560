// voidray_oracle.cpp — CPU restatement of the voidray progressive path-tracing hot path.
//
// *** TEST INFRASTRUCTURE ONLY ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library. The product (libvoidray_cuda.so)
// never links, loads or calls it.
//
// *** PARITY UNPINNED ***  The reference ships no golden vectors, known-answer tests or
// fixtures for this path (its only #[test] is a macro test, voidray_renderer/src/util/vector.rs:36-43)
// and cannot be compiled here (no cargo/rustc; it also needs a live Vulkan queue,
// voidray_renderer/src/render/target.rs:90-131). This file therefore restates the algorithm from
// the reference sources, function by function, citing file:line; its own self-checks
// (tests/test_oracle_*.py) are closed-form cases, not reference outputs.
//
// Third-party arithmetic that is absent from /root/reference (pinned in Cargo.lock) is restated
// from the published algorithms:
//   cgmath 0.18.0 (src/vector.rs, src/structure.rs), for Vector3<f32>:
//     dot(a, b)      = Vector3::mul_element_wise(a, b).sum(), sum() = x + y + z          -> (ax*bx + ay*by) + az*bz
//     cross(a, b)    = (ay*bz - az*by, az*bx - ax*bz, ax*by - ay*bx)
//     magnitude(v)   = sqrt(dot(v, v));  normalize(v) = normalize_to(1) = v * (1 / magnitude(v))
//     angle(a, b)    = Rad::atan2(a.cross(b).magnitude(), dot(a, b))   (the Vector3 specialisation; call site
//                      core/mesh.rs:179 compares `.0 > degrees_to_radians(30.0)` = 30 * PI / 180, util/math.rs:32-34;
//                      known answers around the threshold: tests/test_oracle_geometry.py)
//     Matrix3::new(c0r0, c0r1, c0r2, c1r0, ...) is column-major; Matrix3 * Vector3 = c0 * v.x + c1 * v.y + c2 * v.z
//   rand 0.8.5 (src/distributions/{float,uniform,bernoulli}.rs):
//     f32 from 23 bits   into_float_with_exponent(0): f32::from_bits((u32 >> 9) | 0x3f80_0000) in [1, 2)
//     gen_range(low..high) = UniformFloat::sample_single: scale = high - low; loop { res = (value1_2 - 1.0) * scale + low;
//                          if res < high { return res } }
//     Uniform::new(-1., 1.).sample (rand_distr's samplers): value0_1 * scale + low with scale = 2, low = -1
//     gen::<f32>()       = Standard: (u32 >> 8) as f32 * 2^-24
//     gen_bool(p)        = Bernoulli::new(p): p == 1 -> always true (no draw); p_int = (p * 2^64) as u64; next_u64() < p_int;
//                          BlockRng::next_u64 takes two consecutive u32 words, low word first
//   rand_distr 0.4.3 (src/unit_sphere.rs, unit_disc.rs, unit_circle.rs):
//     UnitSphere  Marsaglia 1972: loop { x1, x2 ~ U(-1, 1); sum = x1*x1 + x2*x2; if sum >= 1 { continue }
//                 factor = 2 * sqrt(1 - sum); return [x1*factor, x2*factor, 1 - 2*sum] }
//     UnitDisc    loop { x1, x2 ~ U(-1, 1); if x1*x1 + x2*x2 <= 1 { return [x1, x2] } }
//     UnitCircle  loop { x1, x2 ~ U(-1, 1); sum = x1*x1 + x2*x2; if sum < 1 { break } };
//                 return [(x1*x1 - x2*x2) / sum, 2*x1*x2 / sum]
//   These are restated from the published crate sources, which are not under /root/reference; the call sites, not the
//   crates, are what the reference pins (iterative.rs:29,39-40, camera.rs:76, simple.rs:77,116,153,220, microfacet.rs:243-266).
//   The reference draws from rand::thread_rng() (OS-seeded ChaCha12, unseedable through the API,
//   voidray_renderer/src/render/iterative.rs:29), so only the *distributions* can be matched. The
//   oracle replaces the generator by Philox4x32-10 keyed by (seed) with counter
//   (pixel, global sample index, block, 0): the draws of one camera sample are a deterministic
//   stream, which is what makes the fixed-seed per-pixel-mean gate possible.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: every f32 operation is a single IEEE
// operation in source order, like rustc without fast-math).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef float F;
const F PI_F = 3.14159265358979323846f;  // std::f32::consts::PI
const F INF_F = std::numeric_limits<float>::infinity();

// ---------------------------------------------------------------------------------------------
// cgmath 0.18.0 restated (Vector3<f32>)
// ---------------------------------------------------------------------------------------------
struct V3 {
    F x, y, z;
};
struct V2 {
    F x, y;
};
inline V3 v3(F x, F y, F z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, F s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(F s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, F s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 mul_elem(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline F dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline F magnitude2(V3 a) { return dot(a, a); }
inline F magnitude(V3 a) { return std::sqrt(magnitude2(a)); }
inline V3 normalize(V3 a) { return a * (1.0f / magnitude(a)); }
inline F angle(V3 a, V3 b) { return std::atan2(magnitude(cross(a, b)), dot(a, b)); }
inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
inline V2 operator*(F s, V2 a) { return V2{a.x * s, a.y * s}; }

// Rust `x as usize` for f32: saturating, NaN -> 0.
inline size_t f32_as_usize(F x) {
    if (!(x > 0.0f)) return 0;  // negatives, -0, NaN
    if (x >= 18446744073709551616.0f) return std::numeric_limits<size_t>::max();
    return (size_t)x;
}
// Rust f32::min / f32::max (IEEE minNum/maxNum: NaN loses)
inline F rmin(F a, F b) { return std::fmin(a, b); }
inline F rmax(F a, F b) { return std::fmax(a, b); }
// compiler-rt __powisf2, what Rust's f32::powi lowers to
inline F powi(F a, int b) {
    const bool recip = b < 0;
    F r = 1.0f;
    while (true) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
// f32::total_cmp
inline int32_t total_key(F f) {
    int32_t i;
    std::memcpy(&i, &f, 4);
    i ^= (int32_t)(((uint32_t)(i >> 31)) >> 1);
    return i;
}

// ---------------------------------------------------------------------------------------------
// Counter-based generator (Philox4x32-10, Salmon et al. SC'11) + rand 0.8.5 / rand_distr 0.4.3
// distribution algorithms
// ---------------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    uint32_t key[2];
    uint32_t pixel, sample;
    uint32_t n;        // draws consumed so far
    uint32_t buf[4];
    uint32_t buf_block;
    Rng(uint64_t seed, uint32_t pixel_, uint32_t sample_) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        pixel = pixel_;
        sample = sample_;
        n = 0;
        buf_block = 0xFFFFFFFFu;
    }
    uint32_t next_u32() {
        const uint32_t block = n >> 2;
        if (block != buf_block) {
            const uint32_t ctr[4] = {pixel, sample, block, 0u};
            philox4x32_10(ctr, key, buf);
            buf_block = block;
        }
        return buf[(n++) & 3u];
    }
    // rand 0.8.5 UniformFloat<f32>: [0,1) with 23 random mantissa bits
    F v01() {
        const uint32_t bits = (next_u32() >> 9) | 0x3f800000u;
        F v12;
        std::memcpy(&v12, &bits, 4);
        return v12 - 1.0f;
    }
    // rng.gen_range(low..high) (UniformFloat::sample_single)
    F gen_range(F low, F high) {
        const F scale = high - low;
        while (true) {
            const F res = v01() * scale + low;
            if (res < high) return res;
        }
    }
    // Uniform::new(-1.0, 1.0).sample(rng): scale = 2, low = -1
    F uniform_m1_1() { return v01() * 2.0f + -1.0f; }
    // rand_distr::UnitSphere
    V3 unit_sphere() {
        while (true) {
            const F x1 = uniform_m1_1();
            const F x2 = uniform_m1_1();
            const F sum = x1 * x1 + x2 * x2;
            if (sum >= 1.0f) continue;
            const F factor = 2.0f * std::sqrt(1.0f - sum);
            return V3{x1 * factor, x2 * factor, 1.0f - 2.0f * sum};
        }
    }
    // rand 0.8.5 Standard f32: 24 random bits * 2^-24
    F gen_f32() { return (F)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    // BlockRng::next_u64: two consecutive words, low word first
    uint64_t next_u64() {
        const uint64_t lo = next_u32();
        const uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }
    // rng.gen_bool(p) = Bernoulli::new(p).sample: p_int = (p * 2^64) as u64, true iff next_u64 < p_int;
    // p == 1 is always true without a draw. (Bernoulli::new panics for p outside [0, 1]; p > 1 is treated as 1.)
    bool gen_bool(double p) {
        if (p >= 1.0) return true;
        const uint64_t p_int = p > 0.0 ? (uint64_t)(p * 18446744073709551616.0) : 0;
        return next_u64() < p_int;
    }
    // rand_distr 0.4.3 UnitCircle
    V2 unit_circle() {
        F x1, x2, sum;
        while (true) {
            x1 = uniform_m1_1();
            x2 = uniform_m1_1();
            sum = x1 * x1 + x2 * x2;
            if (sum < 1.0f) break;
        }
        const F diff = x1 * x1 - x2 * x2;
        return V2{diff / sum, 2.0f * x1 * x2 / sum};
    }
    // rand_distr::UnitDisc
    V2 unit_disc() {
        while (true) {
            const F x1 = uniform_m1_1();
            const F x2 = uniform_m1_1();
            if (x1 * x1 + x2 * x2 <= 1.0f) return V2{x1, x2};
        }
    }
};

// ---------------------------------------------------------------------------------------------
// util/color.rs, util/ray.rs, util/aabb.rs, util/math.rs
// ---------------------------------------------------------------------------------------------
typedef V3 Color;  // util/color.rs:15 `struct Color(pub Vec3)`
const Color BLACK = {0.0f, 0.0f, 0.0f};

// util/color.rs:30-36 — per-channel min only
inline Color color_clamp(Color c, F mx) { return Color{rmin(c.x, mx), rmin(c.y, mx), rmin(c.z, mx)}; }

struct Ray {  // util/ray.rs:5-21
    V3 origin, direction;
    Ray() {}
    Ray(V3 o, V3 d) : origin(o), direction(normalize(d)) {}  // :12-17 normalises
    V3 at(F t) const { return origin + direction * t; }      // :19-21
};

struct HitRecord {  // util/ray.rs:24-49
    V3 point, normal;
    F t;
    V2 uv;
    bool front_face;
    // bookkeeping for the parity gates (not in the reference)
    uint32_t prim;
};
inline HitRecord make_hit(V3 point, V3 outward, F t, V2 uv, const Ray& ray) {  // ray.rs:34-49
    HitRecord h;
    h.front_face = dot(ray.direction, outward) < 0.0f;
    h.normal = h.front_face ? outward : -outward;
    h.point = point;
    h.t = t;
    h.uv = uv;
    h.prim = 0xFFFFFFFFu;
    return h;
}

struct AABB {  // util/aabb.rs:13-19
    V3 min, max;
};
inline AABB aabb_default() { return AABB{v3(INF_F, INF_F, INF_F), v3(-INF_F, -INF_F, -INF_F)}; }  // :151-158
inline bool aabb_eq(const AABB& a, const AABB& b) {
    return a.min.x == b.min.x && a.min.y == b.min.y && a.min.z == b.min.z && a.max.x == b.max.x &&
           a.max.y == b.max.y && a.max.z == b.max.z;
}
inline AABB aabb_point(V3 p) { return AABB{p, p}; }                      // :26-31
inline V3 aabb_centroid(const AABB& b) { return (b.min + b.max) / 2.0f; }  // :33-35
inline AABB aabb_surround(const AABB& a, const AABB& b) {                // :46-59
    return AABB{v3(rmin(a.min.x, b.min.x), rmin(a.min.y, b.min.y), rmin(a.min.z, b.min.z)),
                v3(rmax(a.max.x, b.max.x), rmax(a.max.y, b.max.y), rmax(a.max.z, b.max.z))};
}
inline AABB aabb_epsilon_expand(AABB b, F eps) {  // :62-83
    const V3 dim = v3(b.max.x - b.min.x, b.max.y - b.min.y, b.max.z - b.min.z);
    const V3 c = aabb_centroid(b);
    if (dim.x < eps) { b.min.x = c.x - eps; b.max.x = c.x + eps; }
    if (dim.y < eps) { b.min.y = c.y - eps; b.max.y = c.y + eps; }
    if (dim.z < eps) { b.min.z = c.z - eps; b.max.z = c.z + eps; }
    return b;
}
// util/aabb.rs:86-148
inline bool aabb_hit(const AABB& b, const Ray& ray, F t_min, F t_max) {
    if (aabb_eq(b, aabb_default())) return true;  // :88-90
    {
        const F inv_d = 1.0f / ray.direction.x;
        F t0 = (b.min.x - ray.origin.x) * inv_d;
        F t1 = (b.max.x - ray.origin.x) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    {
        const F inv_d = 1.0f / ray.direction.y;
        F t0 = (b.min.y - ray.origin.y) * inv_d;
        F t1 = (b.max.y - ray.origin.y) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    {
        const F inv_d = 1.0f / ray.direction.z;
        F t0 = (b.min.z - ray.origin.z) * inv_d;
        F t1 = (b.max.z - ray.origin.z) * inv_d;
        if (inv_d < 0.0f) std::swap(t0, t1);
        if (t0 > t_min) t_min = t0;
        if (t1 < t_max) t_max = t1;
        if (t_max <= t_min) return false;
    }
    return true;
}

inline bool near_zero(V3 v) {  // util/math.rs:6-9
    const F EPS = 1.0e-8f;
    return std::fabs(v.x) < EPS && std::fabs(v.y) < EPS && std::fabs(v.z) < EPS;
}
inline V3 reflect(V3 v, V3 n) { return v - 2.0f * dot(v, n) * n; }  // math.rs:12-14
inline V3 refract(V3 uv, V3 n, F etai_over_etat) {                   // math.rs:16-22
    const F cos_theta = rmin(dot(n, -uv), 1.0f);
    const V3 out_perp = etai_over_etat * (uv + cos_theta * n);
    const V3 out_parallel = -std::sqrt(std::fabs(1.0f - magnitude2(out_perp))) * n;
    return out_perp + out_parallel;
}
inline Color lerp_c(Color a, Color b, F t) { return a * (1.0f - t) + b * t; }  // math.rs:41-46

// ---------------------------------------------------------------------------------------------
// Traversal statistics / modes (bookkeeping, not in the reference)
// ---------------------------------------------------------------------------------------------
enum TraverseMode { MODE_FAITHFUL = 0, MODE_EARLY_OUT = 1, MODE_BRUTE = 2 };
struct Counters {
    uint64_t box_tests = 0, tri_tests = 0, segments = 0;
};

// ---------------------------------------------------------------------------------------------
// core/bvh.rs — boxed binary tree, median split, visit both children, right wins ties
// ---------------------------------------------------------------------------------------------
struct BvhNode {
    enum Kind { NONE, OBJECT, SPLIT } kind = NONE;
    size_t handle = 0;
    AABB bounds;
    std::unique_ptr<BvhNode> left, right;
};

struct BoundsCollection {  // bvh.rs:11-15
    virtual ~BoundsCollection() {}
    virtual bool hit(size_t handle, const Ray& ray, F t_min, F t_max, HitRecord& out, Counters& c,
                     int mode) const = 0;
    virtual AABB bounds_ref(size_t handle) const = 0;
};

// bvh.rs:48-130
std::unique_ptr<BvhNode> bvh_from_list(std::vector<size_t>& objects, const BoundsCollection& scene) {
    std::unique_ptr<BvhNode> node(new BvhNode());
    if (objects.empty()) return node;  // BvhNode::None
    if (objects.size() == 1) {
        node->kind = BvhNode::OBJECT;
        node->handle = objects[0];
        return node;
    }
    AABB bounds = scene.bounds_ref(objects[0]);  // reduce(AABB::surround), :57-61
    AABB centroids = aabb_point(aabb_centroid(scene.bounds_ref(objects[0])));  // :64-68
    for (size_t i = 1; i < objects.size(); ++i) {
        const AABB b = scene.bounds_ref(objects[i]);
        bounds = aabb_surround(bounds, b);
        centroids = aabb_surround(centroids, aabb_point(aabb_centroid(b)));
    }
    const V3 spread = centroids.max - centroids.min;  // :70
    int axis;                                          // :71-77
    if (spread.x > spread.y && spread.x > spread.z) axis = 0;
    else if (spread.y > spread.x && spread.y > spread.z) axis = 1;
    else axis = 2;

    // :80-108 — Vec::sort_by is a stable sort; total_cmp ordering
    std::vector<std::pair<int32_t, size_t>> keyed(objects.size());
    for (size_t i = 0; i < objects.size(); ++i) {
        const V3 c = aabb_centroid(scene.bounds_ref(objects[i]));
        keyed[i].first = total_key(axis == 0 ? c.x : (axis == 1 ? c.y : c.z));
        keyed[i].second = objects[i];
    }
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<int32_t, size_t>& a, const std::pair<int32_t, size_t>& b) {
                         return a.first < b.first;
                     });
    for (size_t i = 0; i < objects.size(); ++i) objects[i] = keyed[i].second;

    // :111-120
    const size_t half = objects.size() / 2;
    std::vector<size_t> left_list(objects.begin(), objects.begin() + half);
    std::vector<size_t> right_list(objects.begin() + half, objects.end());
    node->kind = BvhNode::SPLIT;
    node->bounds = bounds;
    node->left = bvh_from_list(left_list, scene);
    node->right = bvh_from_list(right_list, scene);
    return node;
}

// merge_optionals, bvh.rs:165-181 with HitRecord ordering by t (ray.rs:52-62): left only if strictly less
inline bool merge(bool hl, const HitRecord& l, size_t handle_l, bool hr, const HitRecord& r,
                  size_t handle_r, HitRecord& out, size_t& out_handle) {
    if (hl && hr) {
        if (l.t < r.t) { out = l; out_handle = handle_l; }
        else { out = r; out_handle = handle_r; }
        return true;
    }
    if (hl) { out = l; out_handle = handle_l; return true; }
    if (hr) { out = r; out_handle = handle_r; return true; }
    return false;
}

// bvh.rs:132-160. MODE_FAITHFUL is the reference. MODE_EARLY_OUT walks the same tree but narrows
// t_max to the closest hit so far (used only to count the builder-independent n_box/n_tri that
// DESIGN.md's algorithmic-byte figure is defined from).
bool bvh_hit(const BvhNode* node, const Ray& ray, F t_min, F t_max, const BoundsCollection& scene,
             HitRecord& out, size_t& out_handle, Counters& c, int mode) {
    switch (node->kind) {
        case BvhNode::OBJECT: {
            if (scene.hit(node->handle, ray, t_min, t_max, out, c, mode)) {
                out_handle = node->handle;
                return true;
            }
            return false;
        }
        case BvhNode::SPLIT: {
            c.box_tests++;
            if (aabb_hit(node->bounds, ray, t_min, t_max)) {
                HitRecord l, r;
                size_t hl_handle = 0, hr_handle = 0;
                const bool hl = bvh_hit(node->left.get(), ray, t_min, t_max, scene, l, hl_handle, c, mode);
                F t_max_r = t_max;
                if (mode == MODE_EARLY_OUT && hl && l.t < t_max_r) t_max_r = l.t;
                bool hr = bvh_hit(node->right.get(), ray, t_min, t_max_r, scene, r, hr_handle, c, mode);
                if (mode == MODE_EARLY_OUT && hr && !(r.t <= t_max_r)) hr = false;
                return merge(hl, l, hl_handle, hr, r, hr_handle, out, out_handle);
            }
            return false;
        }
        default:
            return false;
    }
}

// ---------------------------------------------------------------------------------------------
// core/texture.rs
// ---------------------------------------------------------------------------------------------
struct ImageTexture {
    std::vector<Color> image;
    size_t width, height;
    int sample_type;  // 0 nearest, 1 bilinear
    Color nearest_sample(F x, F y) const {  // texture.rs:52-57
        const size_t xi = std::min(f32_as_usize(x), width - 1);
        const size_t yi = std::min(f32_as_usize(y), height - 1);
        return image[(yi * width + xi) % image.size()];
    }
    Color bilinear_sample(F x, F y) const {  // texture.rs:59-78
        const size_t len = image.size();
        const size_t x0 = std::min(f32_as_usize(x), width - 1);
        const size_t y0 = std::min(f32_as_usize(y), height - 1);
        const F ax = x - (F)x0;
        const F ay = y - (F)y0;
        return lerp_c(lerp_c(image[(y0 * width + x0) % len], image[(y0 * width + x0 + 1) % len], ax),
                      lerp_c(image[((y0 + 1) * width + x0) % len],
                             image[((y0 + 1) * width + x0 + 1) % len], ax),
                      ay);
    }
    Color sample(F u, F v) const {  // texture.rs:82-98
        if (u < 0.0f) u -= std::trunc(u) - 1.0f;
        if (v < 0.0f) v -= std::trunc(v) - 1.0f;
        const F x = std::fmod(u, 1.0f) * (F)width;
        const F y = (1.0f - std::fmod(v, 1.0f)) * (F)height;
        return sample_type == 0 ? nearest_sample(x, y) : bilinear_sample(x, y);
    }
};

// ---------------------------------------------------------------------------------------------
// core/mesh.rs
// ---------------------------------------------------------------------------------------------
struct Vertex {
    V3 position;
    V2 uv;
    V3 normal;
};
struct Triangle {
    uint32_t vertices[3];
    V3 normal;
};

struct Mesh : BoundsCollection {
    std::vector<Vertex> vertices;
    std::vector<Triangle> triangles;
    AABB bounds;
    std::unique_ptr<BvhNode> bvh_root;  // kind NONE when small

    // mesh.rs:76-116
    void from_buffers(const float* pos, const float* uv, const float* nrm, uint32_t n_vertices,
                      const uint32_t* indices, uint32_t n_indices) {
        vertices.resize(n_vertices);
        for (uint32_t i = 0; i < n_vertices; ++i) {
            vertices[i].position = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
            vertices[i].uv = V2{uv[2 * i], uv[2 * i + 1]};
            vertices[i].normal = v3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        }
        for (uint32_t k = 0; k + 3 <= n_indices; k += 3) {  // chunks_exact(3)
            const V3 e1 = vertices[indices[k]].position - vertices[indices[k + 1]].position;
            const V3 e2 = vertices[indices[k + 2]].position - vertices[indices[k + 1]].position;
            Triangle t;
            t.vertices[0] = indices[k];
            t.vertices[1] = indices[k + 1];
            t.vertices[2] = indices[k + 2];
            t.normal = normalize(cross(e2, e1));
            triangles.push_back(t);
        }
        bounds = aabb_default();
        for (const Vertex& v : vertices) bounds = aabb_surround(bounds, AABB{v.position, v.position});
        bvh_root.reset(new BvhNode());
        if (n_indices > 4 * 3) {  // SMALL_MESH, mesh.rs:43,112
            std::vector<size_t> objs(triangles.size());
            for (size_t i = 0; i < objs.size(); ++i) objs[i] = i;
            bvh_root = bvh_from_list(objs, *this);
        }
    }

    // mesh.rs:193-205
    AABB bounds_ref(size_t handle) const override {
        const Triangle& t = triangles[handle];
        return aabb_epsilon_expand(
            aabb_surround(aabb_point(vertices[t.vertices[0]].position),
                          aabb_surround(aabb_point(vertices[t.vertices[1]].position),
                                        aabb_point(vertices[t.vertices[2]].position))),
            0.001f);
    }

    // Triangle::hit, mesh.rs:144-189 (t_max is ignored by the reference; MODE_EARLY_OUT applies it)
    bool hit(size_t handle, const Ray& ray, F tmin, F tmax, HitRecord& out, Counters& c,
             int mode) const override {
        c.tri_tests++;
        const Triangle& tri = triangles[handle];
        const Vertex& v0 = vertices[tri.vertices[0]];
        const Vertex& v1 = vertices[tri.vertices[1]];
        const Vertex& v2 = vertices[tri.vertices[2]];
        const V3 e1 = v1.position - v0.position;
        const V3 e2 = v2.position - v0.position;
        const V3 h = cross(ray.direction, e2);
        const F a = dot(e1, h);
        if (a > -tmin && a < tmin) return false;
        const F f = 1.0f / a;
        const V3 s = ray.origin - v0.position;
        const F u = f * dot(s, h);
        if (u < 0.0f || u > 1.0f) return false;
        const V3 q = cross(s, e1);
        const F v = f * dot(ray.direction, q);
        if (v < 0.0f || u + v > 1.0f) return false;
        const F t = f * dot(e2, q);
        V3 normal = u * v1.normal + v * v2.normal + (1.0f - u - v) * v0.normal;
        const V2 uv = u * v1.uv + v * v2.uv + (1.0f - u - v) * v0.uv;
        if (angle(normal, tri.normal) > 30.0f * PI_F / 180.0f) normal = tri.normal;  // math.rs:32-34
        if (t > tmin) {
            if (mode == MODE_EARLY_OUT && !(t <= tmax)) return false;
            out = make_hit(ray.at(t), normal, t, uv, ray);
            out.prim = (uint32_t)handle;
            return true;
        }
        return false;
    }

    // Mesh::hit, mesh.rs:122-141
    bool mesh_hit(const Ray& ray, F t_min, F t_max, HitRecord& out, Counters& c, int mode) const {
        if (mode == MODE_BRUTE) {
            // not in the reference: exhaustive closest hit; ties resolved like the reference would
            // if every box test passed (see tie_rank)
            bool any = false;
            for (size_t i = 0; i < triangles.size(); ++i) {
                HitRecord h;
                if (hit(i, ray, t_min, t_max, h, c, MODE_FAITHFUL)) {
                    if (!any || h.t < out.t || (h.t == out.t && tie_rank[i] > tie_rank[out.prim])) out = h;
                    any = true;
                }
            }
            return any;
        }
        if (bvh_root->kind != BvhNode::NONE) {
            size_t handle;
            return bvh_hit(bvh_root.get(), ray, t_min, t_max, *this, out, handle, c, mode);
        }
        bool any = false;
        F closest_so_far = t_max;
        for (size_t i = 0; i < triangles.size(); ++i) {
            HitRecord h;
            if (hit(i, ray, t_min, closest_so_far, h, c, MODE_FAITHFUL)) {
                if (closest_so_far > h.t) {
                    closest_so_far = h.t;
                    out = h;
                    any = true;
                }
            }
        }
        return any;
    }

    // In-order leaf rank of every triangle in the reference tree: among hits with equal t the
    // reference returns the one with the largest rank ("right wins ties", bvh.rs:171); for the
    // small-mesh linear loop the first index wins (mesh.rs:131), so ranks descend with the index.
    std::vector<uint32_t> tie_rank;
    void build_tie_rank() {
        tie_rank.assign(triangles.size(), 0);
        uint32_t next = 0;
        if (bvh_root->kind == BvhNode::NONE) {
            for (size_t i = 0; i < triangles.size(); ++i) tie_rank[i] = (uint32_t)(triangles.size() - 1 - i);
            return;
        }
        std::vector<const BvhNode*> stack;
        stack.push_back(bvh_root.get());
        while (!stack.empty()) {  // in-order == left-to-right leaf order
            const BvhNode* n = stack.back();
            stack.pop_back();
            if (n->kind == BvhNode::OBJECT) tie_rank[n->handle] = next++;
            else if (n->kind == BvhNode::SPLIT) {
                stack.push_back(n->right.get());
                stack.push_back(n->left.get());
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// voidray_common/src/surfaces.rs
// ---------------------------------------------------------------------------------------------
struct AnalyticSurface {
    virtual ~AnalyticSurface() {}
    virtual bool hit(const Ray& ray, F t_min, F t_max, HitRecord& out) const = 0;
    virtual AABB bounds() const = 0;
};
struct Sphere : AnalyticSurface {  // surfaces.rs:31-80
    V3 center;
    F radius;
    AABB bounds() const override {
        return AABB{center - v3(radius, radius, radius), center + v3(radius, radius, radius)};
    }
    bool hit(const Ray& ray, F t_min, F t_max, HitRecord& out) const override {
        const V3 oc = ray.origin - center;
        const F a = magnitude2(ray.direction);
        const F half_b = dot(oc, ray.direction);
        const F c = magnitude2(oc) - radius * radius;
        const F discriminant = half_b * half_b - a * c;
        if (discriminant < 0.0f) return false;
        const F sqrtd = std::sqrt(discriminant);
        F root = (-half_b - sqrtd) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < t_min || t_max < root) return false;
        }
        const V3 point = ray.at(root);
        const V3 normal = (point - center) / radius;
        out = make_hit(point, normal, root, V2{0.0f, 0.0f}, ray);
        return true;
    }
};
struct GroundPlane : AnalyticSurface {  // surfaces.rs:82-114
    F height;
    bool hit(const Ray& ray, F t_min, F t_max, HitRecord& out) const override {
        const F t = (height - ray.origin.y) / ray.direction.y;
        if (t <= t_min || t >= t_max) return false;
        const V3 world_pos = ray.at(t);
        out = make_hit(world_pos, v3(0.0f, 1.0f, 0.0f), t, V2{world_pos.x, world_pos.z}, ray);
        return true;
    }
    AABB bounds() const override {
        return AABB{v3(-INF_F, height - 0.0001f, -INF_F), v3(INF_F, height + 0.0001f, INF_F)};
    }
};

// ---------------------------------------------------------------------------------------------
// Scene (core/scene.rs), materials (voidray_common/src/simple.rs), environments
// ---------------------------------------------------------------------------------------------
struct Scene;

struct Material {  // core/traits.rs:42-50
    virtual ~Material() {}
    // returns attenuation; has_scattered=false terminates the path
    virtual Color scatter(const Scene& scene, const Ray& ray, const HitRecord& hit, Rng& rng,
                          bool& has_scattered, Ray& scattered) const = 0;
};

struct Environment {  // core/traits.rs:66-69
    virtual ~Environment() {}
    virtual Color sample(const Ray& ray) const = 0;
};
struct UniformEnvironment : Environment {  // environments.rs:19-33
    Color color;
    Color sample(const Ray&) const override { return color; }
};
struct HDRIEnvironment : Environment {  // environments.rs:35-86
    std::vector<Color> image;
    size_t width, height;

    // ---- integrator 1 ("fast") only; not in the reference: luminance x sin(theta) sampling tables ----
    // marginal[j] = P(row < j), cond[j * (W + 1) + i] = P(col < i | row j). A texel's weight is
    // (luminance + 5 % of the mean luminance) * sin(polar angle of the row centre); the polar angle follows the
    // lookup's own mapping, environments.rs:80-86: y = (H - 1) - acos(-d.y) / pi * H.
    std::vector<float> marginal, cond;
    void build_sampling_tables() {
        const size_t W = width, H = height;
        std::vector<double> lum(W * H), sinw(H);
        double mean = 0.0;
        for (size_t k = 0; k < W * H; ++k) {
            lum[k] = 0.2126 * (double)image[k].x + 0.7152 * (double)image[k].y + 0.0722 * (double)image[k].z;
            mean += lum[k];
        }
        mean /= (double)(W * H);
        for (size_t j = 0; j < H; ++j) {
            const double sx = 3.14159265358979323846 * ((double)(H - 1) - ((double)j + 0.5)) / (double)H;
            sinw[j] = sx > 0.0 ? std::sin(sx) : 0.0;
        }
        std::vector<double> row_sum(H);
        double total = 0.0;
        for (size_t j = 0; j < H; ++j) {
            double rs = 0.0;
            for (size_t i = 0; i < W; ++i) rs += (lum[j * W + i] + 0.05 * mean) * sinw[j];
            row_sum[j] = rs;
            total += rs;
        }
        marginal.assign(H + 1, 0.0f);
        cond.assign(H * (W + 1), 0.0f);
        double acc = 0.0;
        for (size_t j = 0; j < H; ++j) {
            marginal[j] = (float)(total > 0.0 ? acc / total : (double)j / (double)H);
            acc += row_sum[j];
            double c = 0.0;
            for (size_t i = 0; i < W; ++i) {
                cond[j * (W + 1) + i] = (float)(row_sum[j] > 0.0 ? c / row_sum[j] : (double)i / (double)W);
                c += (lum[j * W + i] + 0.05 * mean) * sinw[j];
            }
            cond[j * (W + 1) + W] = 1.0f;
        }
        marginal[H] = 1.0f;
    }
    // largest k in [0, n) with cdf[k] <= xi (cdf has n + 1 entries, cdf[0] = 0, cdf[n] = 1)
    static size_t cdf_find(const float* cdf, size_t n, F xi) {
        size_t lo = 0, hi = n;
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (cdf[mid] <= xi) lo = mid; else hi = mid;
        }
        return lo;
    }
    V3 sample_direction(Rng& rng) const {
        const size_t W = width, H = height;
        const size_t j = cdf_find(marginal.data(), H, rng.v01());
        const size_t i = cdf_find(cond.data() + j * (W + 1), W, rng.v01());
        const F x = (F)i + rng.v01();
        const F y = (F)j + rng.v01();
        F sx = ((F)(H - 1) - y) / (F)H * PI_F;
        if (!(sx > 1.0e-6f)) sx = 1.0e-6f;
        const F ang = x / (F)W * (2.0f * PI_F) - PI_F;
        const F r = std::sin(sx);
        return v3(r * std::cos(ang), -std::cos(sx), -(r * std::sin(ang)));
    }
    F pdf_direction(V3 dir) const {
        const size_t W = width, H = height;
        const V3 d = normalize(dir);
        const F sx = std::acos(-d.y);
        const F sy = std::atan2(-d.z, d.x) + PI_F;
        const F x = sy / (2.0f * PI_F) * (F)W;
        const F y = (F)(H - 1) - (sx / PI_F * (F)H);
        if (!(y >= 0.0f)) return 0.0f;
        const size_t i = std::min(f32_as_usize(x), W - 1);
        const size_t j = std::min(f32_as_usize(y), H - 1);
        const F p_tex = (marginal[j + 1] - marginal[j]) * (cond[j * (W + 1) + i + 1] - cond[j * (W + 1) + i]);
        return p_tex * ((F)W * (F)H) / (2.0f * PI_F * PI_F * rmax(std::sin(sx), 1.0e-6f));
    }
    Color bilinear_sample(F x, F y) const {  // :57-76
        const size_t len = image.size();
        const size_t x0 = std::min(f32_as_usize(x), width - 1);
        const size_t y0 = std::min(f32_as_usize(y), height - 1);
        const F ax = x - (F)x0;
        const F ay = y - (F)y0;
        return lerp_c(lerp_c(image[(y0 * width + x0) % len], image[(y0 * width + x0 + 1) % len], ax),
                      lerp_c(image[((y0 + 1) * width + x0) % len],
                             image[((y0 + 1) * width + x0 + 1) % len], ax),
                      ay);
    }
    Color sample(const Ray& ray) const override {  // :80-86 ; math.rs:24-29
        const V3 d = normalize(ray.direction);
        const F sx = std::acos(-d.y);
        const F sy = std::atan2(-d.z, d.x) + PI_F;
        const F u = sx / PI_F;
        const F v = sy / (2.0f * PI_F);
        const F x = v * (F)width;
        const F y = (F)(height - 1) - (u * (F)height);
        return bilinear_sample(x, y);
    }
};

struct Camera {  // core/camera.rs:7-22 (public fields) + CameraAcceleration :57-66
    V3 eye, direction, up;
    F fov;
    bool has_dof;
    F aperture;
    V3 focal_point;
    // acceleration
    V3 right;
    F d;
    F focal_length;
    void build_acceleration() {  // camera.rs:38-54
        d = 1.0f / std::tan(fov / 2.0f);
        right = normalize(cross(direction, up));
        focal_length = has_dof ? dot(focal_point - eye, direction) : 0.0f;
    }
    Ray cast_ray(F x, F y, Rng& rng) const {  // camera.rs:69-82
        V3 origin = eye;
        V3 new_dir = d * direction + x * right + y * up;
        if (has_dof) {
            const V3 fp = origin + normalize(new_dir) * focal_length;
            const V2 s = rng.unit_disc();
            origin = origin + (s.x * right + s.y * up) * aperture;
            new_dir = fp - origin;
        }
        return Ray(origin, normalize(new_dir));
    }
};

struct Surface {  // core/traits.rs:59-63
    bool is_mesh;
    size_t mesh;
    std::shared_ptr<AnalyticSurface> analytic;
};
struct Object {  // core/scene.rs:31-34
    size_t surface, material;
};

struct Scene : BoundsCollection {
    Camera camera;
    std::vector<ImageTexture> textures;
    std::vector<Object> objects;
    std::vector<Surface> surfaces;
    std::vector<std::shared_ptr<Mesh>> meshes;
    std::vector<std::shared_ptr<Material>> materials;
    std::shared_ptr<Environment> environment;
    std::unique_ptr<BvhNode> bvh;
    std::string error;

    // scene.rs:72-92
    AABB bounds_ref(size_t handle) const override {
        const Surface& s = surfaces[handle];
        return s.is_mesh ? meshes[s.mesh]->bounds : s.analytic->bounds();
    }
    bool hit(size_t handle, const Ray& ray, F t_min, F t_max, HitRecord& out, Counters& c,
             int mode) const override {
        const Surface& s = surfaces[handle];
        if (s.is_mesh) return meshes[s.mesh]->mesh_hit(ray, t_min, t_max, out, c, mode);
        return s.analytic->hit(ray, t_min, t_max, out);
    }
    // scene.rs:163-179
    void build_acceleration() {
        camera.build_acceleration();
        std::vector<size_t> objs(surfaces.size());
        for (size_t i = 0; i < objs.size(); ++i) objs[i] = i;
        bvh = bvh_from_list(objs, *this);
        for (auto& m : meshes) m->build_tie_rank();
        // scene-level in-order surface rank (bookkeeping for MODE_BRUTE ties between surfaces)
        surface_rank.assign(surfaces.size(), 0);
        uint32_t next = 0;
        std::vector<const BvhNode*> stack;
        stack.push_back(bvh.get());
        while (!stack.empty()) {
            const BvhNode* n = stack.back();
            stack.pop_back();
            if (n->kind == BvhNode::OBJECT) surface_rank[n->handle] = next++;
            else if (n->kind == BvhNode::SPLIT) {
                stack.push_back(n->right.get());
                stack.push_back(n->left.get());
            }
        }
    }
    std::vector<uint32_t> surface_rank;

    // SceneAcceleration::hit, scene.rs:182-185 — t_min = 1e-5, t_max = INF. Returns the *surface*
    // handle, which the reference then uses as an *object* index.
    bool scene_hit(const Ray& ray, HitRecord& out, size_t& surface, Counters& c, int mode) const {
        c.segments++;
        if (mode == MODE_BRUTE) {
            bool any = false;
            for (size_t s = 0; s < surfaces.size(); ++s) {
                HitRecord h;
                if (hit(s, ray, 0.00001f, INF_F, h, c, mode)) {
                    if (!any || h.t < out.t || (h.t == out.t && surface_rank[s] > surface_rank[surface])) {
                        out = h;
                        surface = s;
                    }
                    any = true;
                }
            }
            return any;
        }
        return bvh_hit(bvh.get(), ray, 0.00001f, INF_F, *this, out, surface, c, mode);
    }
};

// simple.rs:88-132
struct Lambertian : Material {
    bool albedo_is_texture;
    Color albedo;
    size_t albedo_tex;
    bool has_normal_tex;
    size_t normal_tex;
    Color scatter(const Scene& scene, const Ray&, const HitRecord& hit, Rng& rng, bool& has_scattered,
                  Ray& scattered) const override {
        const V3 normal = has_normal_tex ? scene.textures[normal_tex].sample(hit.uv.x, hit.uv.y) : hit.normal;
        V3 scatter_direction = normal + rng.unit_sphere();
        if (near_zero(scatter_direction)) scatter_direction = normal;
        scattered = Ray(hit.point, scatter_direction);
        has_scattered = true;
        return albedo_is_texture ? scene.textures[albedo_tex].sample(hit.uv.x, hit.uv.y) : albedo;
    }
};
// simple.rs:135-160
struct Metal : Material {
    Color albedo;
    F fuzz;
    Color scatter(const Scene&, const Ray& ray, const HitRecord& hit, Rng& rng, bool& has_scattered,
                  Ray& scattered) const override {
        const V3 reflected = normalize(reflect(ray.direction, hit.normal));
        // The reference loop is unbounded (simple.rs:150-158). A path whose reflected direction lies
        // below the shading hemisphere with fuzz < 1 would spin forever; the oracle (and the CUDA
        // path, identically) gives up after 64 rejected draws and terminates the path with BLACK
        // (no light is injected where the reference would never return). Documented deviation,
        // DESIGN.md "Reference quirks".
        for (int i = 0; i < 64; ++i) {
            scattered = Ray(hit.point, reflected + fuzz * rng.unit_sphere());
            if (dot(scattered.direction, hit.normal) > 0.0f) {
                has_scattered = true;
                return albedo;
            }
        }
        has_scattered = false;
        return Color{0.0f, 0.0f, 0.0f};
    }
};
// simple.rs:163-184
struct Emission : Material {
    Color color;  // already multiplied by strength
    Color scatter(const Scene&, const Ray&, const HitRecord&, Rng&, bool& has_scattered, Ray&) const override {
        has_scattered = false;
        return color;
    }
};
// simple.rs:187-231
struct Dielectric : Material {
    F ir;
    static F reflectance(F cosine, F idx) {  // :192-197
        F r0 = (1.0f - idx) / (1.0f + idx);
        r0 = r0 * r0;
        return r0 + (1.0f - r0) * powi(1.0f - cosine, 5);
    }
    Color scatter(const Scene&, const Ray& ray, const HitRecord& hit, Rng& rng, bool& has_scattered,
                  Ray& scattered) const override {
        const F refraction_ratio = hit.front_face ? 1.0f / ir : ir;
        const V3 unit_direction = normalize(ray.direction);
        const F cos_theta = rmin(dot(hit.normal, -unit_direction), 1.0f);
        const F sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
        const bool cannot_refract = (refraction_ratio * sin_theta) > 1.0f;
        V3 direction;
        // `||` short-circuits: no draw is consumed under total internal reflection
        if (cannot_refract || reflectance(cos_theta, refraction_ratio) > rng.gen_range(0.0f, 1.0f))
            direction = reflect(unit_direction, hit.normal);
        else
            direction = refract(unit_direction, hit.normal, refraction_ratio);
        scattered = Ray(hit.point, direction);
        has_scattered = true;
        return Color{1.0f, 1.0f, 1.0f};
    }
};
// simple.rs:60-81 through the blanket impl core/traits.rs:23-40
struct LambertianBSDF : Material {
    Color albedo;
    Color scatter(const Scene&, const Ray& ray, const HitRecord& hit, Rng& rng, bool& has_scattered,
                  Ray& scattered) const override {
        const V3 wo = normalize(ray.direction);
        (void)wo;
        const V3 wi = normalize(rng.unit_sphere());  // sample(): pdf 1.0 (sic)
        const F pdf = 1.0f;
        const Color f = albedo / PI_F;
        scattered = Ray(ray.at(hit.t), wi);
        has_scattered = true;
        return f * std::fabs(dot(wi, hit.normal)) * (1.0f / pdf);
    }
};

// util/math.rs:50-60 — note Matrix3::new is column-major, so the product below evaluates
// (ns . h, nss . h, normal . h): the transpose of a local-to-world basis. Reproduced as written.
struct M3 {
    V3 c0, c1, c2;
};
inline V3 m3_mul(const M3& m, V3 v) { return m.c0 * v.x + m.c1 * v.y + m.c2 * v.z; }
inline M3 local_to_world(V3 normal) {
    const V3 ns = std::isnormal(normal.x) ? normalize(v3(normal.y, -normal.x, 0.0f)) : normalize(v3(0.0f, -normal.z, normal.y));
    const V3 nss = cross(normal, ns);
    return M3{v3(ns.x, nss.x, normal.x), v3(ns.y, nss.y, normal.y), v3(ns.z, nss.z, normal.z)};
}
inline V3 lerp_v(V3 a, V3 b, F t) { return a * (1.0f - t) + b * t; }
inline bool sign_positive(F x) { return !std::signbit(x); }
inline F signum(F x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0f : 1.0f); }

// voidray_common/src/microfacet.rs through the blanket impl core/traits.rs:23-40
struct MicrofacetBSDF : Material {
    Color color;
    F index, roughness, metallic, emittance;
    bool transparent;

    // microfacet.rs:120-208
    Color bsdf(V3 n, V3 wo, V3 wi) const {
        const F n_dot_wi = dot(n, wi);
        const F n_dot_wo = dot(n, wo);
        const bool wi_outside = sign_positive(n_dot_wi);
        const bool wo_outside = sign_positive(n_dot_wo);
        if (!transparent && (!wi_outside || !wo_outside)) return BLACK;
        if (wi_outside == wo_outside) {
            const V3 h = normalize(wi + wo);
            const F wo_dot_h = dot(wo, h);
            const F n_dot_h = dot(n, h);
            const F nh2 = powi(n_dot_h, 2);
            const F m2 = roughness * roughness;
            const F d = std::exp((nh2 - 1.0f) / (m2 * nh2)) / (m2 * PI_F * nh2 * nh2);
            V3 f;
            if (!wi_outside && std::sqrt(1.0f - wo_dot_h * wo_dot_h) * index > 1.0f) {
                f = v3(1.0f, 1.0f, 1.0f);
            } else {
                const F f0s = powi((index - 1.0f) / (index + 1.0f), 2);
                const V3 f0 = lerp_v(v3(f0s, f0s, f0s), color, metallic);
                f = f0 + (v3(1.0f, 1.0f, 1.0f) - f0) * powi(1.0f - wo_dot_h, 5);
            }
            F g = rmin(n_dot_wi * n_dot_h, n_dot_wo * n_dot_h);
            g = (2.0f * g) / wo_dot_h;
            g = rmin(g, 1.0f);
            const V3 specular = d * f * g / (4.0f * n_dot_wo * n_dot_wi);
            if (transparent) return specular;
            const V3 diffuse = mul_elem(v3(1.0f, 1.0f, 1.0f) - f, color) / PI_F;
            return specular + diffuse;
        }
        const F eta_t = wo_outside ? index : 1.0f / index;
        const V3 h = normalize(wi * eta_t + wo);
        const F wi_dot_h = dot(wi, h);
        const F wo_dot_h = dot(wo, h);
        const F n_dot_h = dot(n, h);
        const F nh2 = powi(n_dot_h, 2);
        const F m2 = roughness * roughness;
        const F d = std::exp((nh2 - 1.0f) / (m2 * nh2)) / (m2 * PI_F * nh2 * nh2);
        const F f0s = powi((index - 1.0f) / (index + 1.0f), 2);
        const V3 f0 = lerp_v(v3(f0s, f0s, f0s), color, metallic);
        const V3 f = f0 + (v3(1.0f, 1.0f, 1.0f) - f0) * powi(1.0f - std::fabs(wi_dot_h), 5);
        F g = rmin(std::fabs(n_dot_wi * n_dot_h), std::fabs(n_dot_wo * n_dot_h));
        g = (2.0f * g) / std::fabs(wo_dot_h);
        g = rmin(g, 1.0f);
        const V3 btdf = std::fabs(wi_dot_h * wo_dot_h / (n_dot_wi * n_dot_wo)) *
                        (d * (v3(1.0f, 1.0f, 1.0f) - f) * g / powi(eta_t * wi_dot_h + wo_dot_h, 2));
        return mul_elem(btdf, color);
    }

    V3 beckmann(V3 n, F m2, Rng& rng) const {  // :239-249
        const F theta = std::atan(std::sqrt(m2 * -std::log(rng.gen_f32())));
        const F sin_t = std::sin(theta), cos_t = std::cos(theta);
        const V2 c = rng.unit_circle();
        return m3_mul(local_to_world(n), v3(c.x * sin_t, c.y * sin_t, cos_t));
    }
    F beckmann_pdf(V3 h, V3 n, F m2) const {  // :251-256
        const F cos_t = std::fabs(dot(h, n));
        const F sin_t = std::sqrt(1.0f - cos_t * cos_t);
        return (1.0f / (PI_F * m2 * powi(cos_t, 3))) * std::exp(-powi(sin_t / cos_t, 2) / m2);
    }

    // microfacet.rs:222-313; returns false for None
    bool sample(V3 n, V3 wo, Rng& rng, V3& wi_out, F& pdf_out) const {
        const F m2 = roughness * roughness;
        const F f0 = powi((index - 1.0f) / (index + 1.0f), 2);
        F f = (1.0f - metallic) * f0 + metallic * ((color.x + color.y + color.z) / 3.0f);
        f = f * (1.0f - 0.2f) + 1.0f * 0.2f;  // lerp(f, 1.0, 0.2)
        const F eta_t = dot(wo, n) > 0.0f ? index : 1.0f / index;
        V3 wi;
        if (rng.gen_bool((double)f)) {
            const V3 h = beckmann(n, m2, rng);
            wi = -reflect(wo, h);
        } else if (!transparent) {
            const V2 dsk = rng.unit_disc();
            const F z = std::sqrt(1.0f - dsk.x * dsk.x - dsk.y * dsk.y);
            wi = m3_mul(local_to_world(n), v3(dsk.x, dsk.y, z));
        } else {
            const V3 h = beckmann(n, m2, rng);
            const F cos_to = dot(h, wo);
            const V3 wo_perp = wo - h * cos_to;
            const V3 wi_perp = -wo_perp / eta_t;
            const F sin2_ti = magnitude2(wi_perp);
            if (sin2_ti > 1.0f) return false;
            const F cos_ti = std::sqrt(1.0f - sin2_ti);
            wi = -signum(cos_to) * cos_ti * h + wi_perp;
        }
        F p = 0.0f;
        {
            const V3 h = normalize(wi + wo);
            const F p_h = beckmann_pdf(h, n, m2);
            p += f * p_h / (4.0f * std::fabs(dot(h, wo)));
        }
        if (!transparent) {
            p += (1.0f - f) * rmax(dot(wi, n), 0.0f) / PI_F;
        } else if (sign_positive(dot(wo, n)) != sign_positive(dot(wi, n))) {
            const V3 h = normalize(wi * eta_t + wo);
            const F p_h = beckmann_pdf(h, n, m2);
            const F h_dot_wo = dot(h, wo);
            const F h_dot_wi = dot(h, wi);
            const F jacobian = std::fabs(h_dot_wo) / powi(eta_t * h_dot_wi + h_dot_wo, 2);
            p += (1.0f - f) * p_h * jacobian;
        } else {
            p += 0.0f;
        }
        if (p == 0.0f) return false;
        wi_out = wi;
        pdf_out = p;
        return true;
    }

    // the blanket `impl<M: BSDFMaterial> Material for M`, core/traits.rs:23-40 (wo is the *incoming* direction)
    Color scatter(const Scene&, const Ray& ray, const HitRecord& hit, Rng& rng, bool& has_scattered,
                  Ray& scattered) const override {
        const V3 wo = normalize(ray.direction);
        V3 wi;
        F pdf;
        if (sample(hit.normal, wo, rng, wi, pdf)) {
            const Color f = bsdf(hit.normal, wo, wi);
            scattered = Ray(ray.at(hit.t), wi);
            has_scattered = true;
            return f * std::fabs(dot(wi, hit.normal)) * (1.0f / pdf);
        }
        has_scattered = false;
        return BLACK;  // hex_color(0x000000)
    }
};

// ---------------------------------------------------------------------------------------------
// core/tracer.rs
// ---------------------------------------------------------------------------------------------
struct RenderSettings {  // core/settings.rs:15-33 + oracle-only switches
    uint32_t total_samples;
    uint32_t max_bounces;
    F firefly_clamp;
    int32_t render_mode;    // 0 = Full, 1 = Normal
    int32_t pixel_mapping;  // 0 = fixed (y = index / width), 1 = reference (iterative.rs:26,33)
    uint64_t seed;
    int32_t traverse_mode;
    int32_t integrator;  // 0 = the reference estimator; 1 = "fast" (not in the reference)
};

// Density (per solid angle) of the reference's Lambertian direction normalize(n + UnitSphere)
// (simple.rs:116) for an arbitrary, possibly non-unit n: the points n + s lie on the unit sphere around n;
// a ray from the origin along w meets it at r = a c +- sqrt(a^2 c^2 - a^2 + 1) (a = |n|, c = cos(w, n)), and
// projecting the uniform surface measure gives sum over positive roots of r^2 / (4 pi |r - a c|).
// For |n| = 1 this is cos/pi.
inline F lambert_reference_pdf(V3 w_unit, V3 n) {
    const F a = magnitude(n);
    if (!(a > 1.0e-6f)) return 1.0f / (4.0f * PI_F);
    const F ac = dot(w_unit, n);
    const F disc = ac * ac - a * a + 1.0f;
    if (!(disc >= 0.0f)) return 0.0f;
    const F sq = rmax(std::sqrt(disc), 1.0e-6f);
    const F r1 = ac + sq, r2 = ac - sq;
    F sum = 0.0f;
    if (r1 > 0.0f) sum += r1 * r1;
    if (r2 > 0.0f) sum += r2 * r2;
    return sum / (4.0f * PI_F * sq);
}

// integrator 1, Lambertian: same integrand as simple.rs:103-132 (albedo x the density above), sampled by
// one-sample MIS between that density and the HDRI table. Returns the attenuation albedo * p_ref / p_mix.
Color lambertian_fast(const Scene& scene, const Lambertian& m, const HitRecord& hit, Rng& rng, bool& has_scattered,
                      Ray& scattered);

Color trace_ray_internal(const Scene& scene, const RenderSettings& st, const Ray& ray, uint32_t depth,
                         Rng& rng, Counters& c) {  // tracer.rs:19-56
    Color color = BLACK;
    if (depth < st.max_bounces) {
        HitRecord hit;
        size_t surface = 0;
        if (!scene.scene_hit(ray, hit, surface, c, st.traverse_mode)) {
            if (!scene.environment) return BLACK;
            return scene.environment->sample(ray);
        }
        // scene.rs:183-184: the surface index is used as an object index
        const Material& material = *scene.materials[scene.objects[surface].material];
        Color attenuation;
        bool has_scattered = false;
        Ray scattered;
        const Lambertian* lam = st.integrator == 1 ? dynamic_cast<const Lambertian*>(&material) : nullptr;
        if (st.render_mode == 0 && lam) {
            attenuation = lambertian_fast(scene, *lam, hit, rng, has_scattered, scattered);
        } else if (st.render_mode == 0) {
            attenuation = material.scatter(scene, ray, hit, rng, has_scattered, scattered);
        } else {
            attenuation = 0.5f * normalize(hit.normal) + v3(1.0f, 1.0f, 1.0f) * 0.5f;
        }
        // integrator 1: Russian roulette from the fourth segment on. Survival probability = the largest
        // attenuation channel clamped to [0.05, 1]; survivors are divided by it, the rest see BLACK.
        bool killed = false;
        if (st.integrator == 1 && has_scattered && depth >= 3) {
            const F q = rmin(rmax(rmax(attenuation.x, rmax(attenuation.y, attenuation.z)), 0.05f), 1.0f);
            if (rng.v01() < q) attenuation = attenuation / q;
            else killed = true;
        }
        Color delta;
        if (has_scattered && killed) delta = mul_elem(attenuation, BLACK);
        else if (has_scattered) delta = mul_elem(attenuation, trace_ray_internal(scene, st, scattered, depth + 1, rng, c));
        else delta = attenuation;
        color = color + color_clamp(delta, st.firefly_clamp);
    }
    return color;
}

Color lambertian_fast(const Scene& scene, const Lambertian& m, const HitRecord& hit, Rng& rng, bool& has_scattered,
                      Ray& scattered) {
    const V3 normal = m.has_normal_tex ? scene.textures[m.normal_tex].sample(hit.uv.x, hit.uv.y) : hit.normal;
    const Color albedo = m.albedo_is_texture ? scene.textures[m.albedo_tex].sample(hit.uv.x, hit.uv.y) : m.albedo;
    const HDRIEnvironment* env = dynamic_cast<const HDRIEnvironment*>(scene.environment.get());
    V3 w;
    if (env && rng.v01() >= 0.5f) {
        w = normalize(env->sample_direction(rng));
    } else {
        V3 dir = normal + rng.unit_sphere();
        if (near_zero(dir)) dir = normal;
        w = normalize(dir);
    }
    scattered = Ray(hit.point, w);
    const F p_ref = lambert_reference_pdf(scattered.direction, normal);
    const F p_mix = env ? 0.5f * p_ref + 0.5f * env->pdf_direction(scattered.direction) : p_ref;
    if (!(p_mix > 0.0f) || !(p_ref > 0.0f)) {
        has_scattered = false;
        return BLACK;
    }
    has_scattered = true;
    return albedo * (p_ref / p_mix);
}

// render/iterative.rs:25-33 — pixel index -> camera-plane coordinates
inline void pixel_to_ndc(uint32_t index, uint32_t W, uint32_t H, int mapping, F& x, F& y, F& d) {
    const uint32_t px = index % W;
    const uint32_t py = mapping == 1 ? index / H : index / W;
    d = (F)std::max(W, H);
    x = ((F)(2u * px + 1u) - (F)W) / d;
    y = ((F)(2u * (H - py) - 1u) - (F)H) / d;  // u32 wrapping arithmetic, as in a release build
}

// One camera sample: iterative.rs:38-42
inline Ray camera_sample_ray(const Scene& scene, F x, F y, F d, Rng& rng) {
    const F dx = rng.gen_range(-1.0f / d, 1.0f / d);
    const F dy = rng.gen_range(-1.0f / d, 1.0f / d);
    return scene.camera.cast_ray(x + dx, y + dy, rng);
}

}  // namespace

// =============================================================================================
// C interface (ctypes) — mirrors the reference's Scene builder (core/scene.rs:94-161)
// =============================================================================================
extern "C" {

struct vo_material_desc {
    int32_t kind;  // 0 lambertian, 1 metal, 2 dielectric, 3 emission, 4 lambertian_bsdf, 5 microfacet
    float color[3];
    float param;          // metal: fuzz; dielectric: ir; emission: strength
    int32_t albedo_tex;   // -1: use color
    int32_t normal_tex;   // -1: none
    float index, roughness, metallic, emittance;  // microfacet (voidray_common/src/microfacet.rs:9-27)
    int32_t transparent;
};

struct vo_settings {
    uint32_t total_samples;
    uint32_t max_bounces;
    float firefly_clamp;
    int32_t render_mode;
    int32_t pixel_mapping;
    int32_t traverse_mode;
    uint64_t seed;
    int32_t integrator;
    int32_t reserved;
};

struct vo_counters {
    uint64_t box_tests, tri_tests, segments;
};

void* vo_scene_create() {
    Scene* s = new Scene();
    // Scene::empty(), scene.rs:95-111 — Camera::look_at((1,0,10),(0,0,0),(0,1,0), PI/6)
    const V3 eye = v3(1.0f, 0.0f, 10.0f), center = v3(0.0f, 0.0f, 0.0f), up0 = v3(0.0f, 1.0f, 0.0f);
    const V3 direction = normalize(center - eye);
    const V3 up = normalize(up0 - dot(up0, direction) * direction);
    s->camera.eye = eye;
    s->camera.direction = direction;
    s->camera.up = up;
    s->camera.fov = PI_F / 6.0f;
    s->camera.has_dof = false;
    s->camera.aperture = 0.0f;
    s->camera.focal_point = v3(0, 0, 0);
    return s;
}
void vo_scene_destroy(void* p) { delete (Scene*)p; }

uint32_t vo_add_texture(void* p, const float* rgb, uint32_t w, uint32_t h, int32_t sample_type) {
    Scene* s = (Scene*)p;
    ImageTexture t;
    t.width = w;
    t.height = h;
    t.sample_type = sample_type;
    t.image.resize((size_t)w * h);
    for (size_t i = 0; i < t.image.size(); ++i) t.image[i] = v3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    s->textures.push_back(std::move(t));
    return (uint32_t)s->textures.size() - 1;
}
uint32_t vo_add_mesh(void* p, const float* pos, const float* uv, const float* nrm, uint32_t n_vertices,
                     const uint32_t* indices, uint32_t n_indices) {
    Scene* s = (Scene*)p;
    std::shared_ptr<Mesh> m(new Mesh());
    m->from_buffers(pos, uv, nrm, n_vertices, indices, n_indices);
    s->meshes.push_back(m);
    Surface sf;
    sf.is_mesh = true;
    sf.mesh = s->meshes.size() - 1;
    s->surfaces.push_back(sf);
    return (uint32_t)s->surfaces.size() - 1;
}
uint32_t vo_add_sphere(void* p, const float* c, float r) {
    Scene* s = (Scene*)p;
    std::shared_ptr<Sphere> sp(new Sphere());
    sp->center = v3(c[0], c[1], c[2]);
    sp->radius = r;
    Surface sf;
    sf.is_mesh = false;
    sf.mesh = 0;
    sf.analytic = sp;
    s->surfaces.push_back(sf);
    return (uint32_t)s->surfaces.size() - 1;
}
uint32_t vo_add_ground_plane(void* p, float height) {
    Scene* s = (Scene*)p;
    std::shared_ptr<GroundPlane> g(new GroundPlane());
    g->height = height;
    Surface sf;
    sf.is_mesh = false;
    sf.mesh = 0;
    sf.analytic = g;
    s->surfaces.push_back(sf);
    return (uint32_t)s->surfaces.size() - 1;
}
int32_t vo_add_material(void* p, const vo_material_desc* d) {
    Scene* s = (Scene*)p;
    const Color col = v3(d->color[0], d->color[1], d->color[2]);
    std::shared_ptr<Material> m;
    switch (d->kind) {
        case 0: {
            Lambertian* l = new Lambertian();
            l->albedo_is_texture = d->albedo_tex >= 0;
            l->albedo = col;
            l->albedo_tex = d->albedo_tex >= 0 ? (size_t)d->albedo_tex : 0;
            l->has_normal_tex = d->normal_tex >= 0;
            l->normal_tex = d->normal_tex >= 0 ? (size_t)d->normal_tex : 0;
            m.reset(l);
            break;
        }
        case 1: {
            Metal* t = new Metal();
            t->albedo = col;
            t->fuzz = d->param;
            m.reset(t);
            break;
        }
        case 2: {
            Dielectric* t = new Dielectric();
            t->ir = d->param;
            m.reset(t);
            break;
        }
        case 3: {
            Emission* t = new Emission();
            t->color = col * d->param;  // Emission::new, simple.rs:168-172
            m.reset(t);
            break;
        }
        case 4: {
            LambertianBSDF* t = new LambertianBSDF();
            t->albedo = col;
            m.reset(t);
            break;
        }
        case 5: {
            MicrofacetBSDF* t = new MicrofacetBSDF();
            t->color = col;
            t->index = d->index;
            t->roughness = d->roughness;
            t->metallic = d->metallic;
            t->emittance = d->emittance;
            t->transparent = d->transparent != 0;
            m.reset(t);
            break;
        }
        default:
            return -1;
    }
    s->materials.push_back(m);
    return (int32_t)s->materials.size() - 1;
}
uint32_t vo_add_object(void* p, uint32_t material, uint32_t surface) {
    Scene* s = (Scene*)p;
    s->objects.push_back(Object{surface, material});
    return (uint32_t)s->objects.size() - 1;
}
void vo_set_camera(void* p, const float* eye, const float* direction, const float* up, float fov,
                   int32_t has_dof, float aperture, const float* focal_point) {
    Scene* s = (Scene*)p;
    s->camera.eye = v3(eye[0], eye[1], eye[2]);
    s->camera.direction = v3(direction[0], direction[1], direction[2]);
    s->camera.up = v3(up[0], up[1], up[2]);
    s->camera.fov = fov;
    s->camera.has_dof = has_dof != 0;
    s->camera.aperture = aperture;
    if (focal_point) s->camera.focal_point = v3(focal_point[0], focal_point[1], focal_point[2]);
}
// Camera::look_at, camera.rs:26-36 — returns direction[3], up[3]
void vo_look_at(const float* eye, const float* center, const float* up_in, float* direction, float* up) {
    const V3 e = v3(eye[0], eye[1], eye[2]), c = v3(center[0], center[1], center[2]);
    const V3 u0 = v3(up_in[0], up_in[1], up_in[2]);
    const V3 dir = normalize(c - e);
    const V3 u = normalize(u0 - dot(u0, dir) * dir);
    direction[0] = dir.x; direction[1] = dir.y; direction[2] = dir.z;
    up[0] = u.x; up[1] = u.y; up[2] = u.z;
}
void vo_set_environment_uniform(void* p, const float* rgb) {
    Scene* s = (Scene*)p;
    std::shared_ptr<UniformEnvironment> e(new UniformEnvironment());
    e->color = v3(rgb[0], rgb[1], rgb[2]);
    s->environment = e;
}
void vo_set_environment_hdri(void* p, const float* rgb, uint32_t w, uint32_t h) {
    Scene* s = (Scene*)p;
    std::shared_ptr<HDRIEnvironment> e(new HDRIEnvironment());
    e->width = w;
    e->height = h;
    e->image.resize((size_t)w * h);
    for (size_t i = 0; i < e->image.size(); ++i) e->image[i] = v3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    e->build_sampling_tables();
    s->environment = e;
}
void vo_clear_environment(void* p) { ((Scene*)p)->environment.reset(); }

// build_acceleration. Returns 0, or -1 if a surface has no object at its own index (the reference
// would panic on the first hit of that surface, scene.rs:183-184 + :133-135).
int32_t vo_commit(void* p) {
    Scene* s = (Scene*)p;
    s->build_acceleration();
    if (s->objects.size() < s->surfaces.size()) return -1;
    for (const Object& o : s->objects)
        if (o.material >= s->materials.size()) return -1;
    return 0;
}

// Global tie rank of every triangle of a mesh surface, in the reference trees' in-order sequence.
void vo_mesh_tie_rank(void* p, uint32_t surface, uint32_t* out) {
    Scene* s = (Scene*)p;
    const Mesh& m = *s->meshes[s->surfaces[surface].mesh];
    for (size_t i = 0; i < m.tie_rank.size(); ++i) out[i] = m.tie_rank[i];
}
void vo_surface_rank(void* p, uint32_t* out) {
    Scene* s = (Scene*)p;
    for (size_t i = 0; i < s->surface_rank.size(); ++i) out[i] = s->surface_rank[i];
}

static RenderSettings to_settings(const vo_settings* v) {
    RenderSettings st;
    st.total_samples = v->total_samples;
    st.max_bounces = v->max_bounces;
    st.firefly_clamp = v->firefly_clamp;
    st.render_mode = v->render_mode;
    st.pixel_mapping = v->pixel_mapping;
    st.seed = v->seed;
    st.traverse_mode = v->traverse_mode;
    st.integrator = v->integrator;
    return st;
}

// Closest hit of arbitrary rays (direction is normalised by Ray::new like every reference ray).
// surface = 0xFFFFFFFF on a miss. normal/uv/front may be null.
void vo_trace_rays(void* p, uint64_t n, const float* origins, const float* dirs, int32_t mode,
                   uint32_t* surface, uint32_t* prim, float* t, float* normal, float* uv, uint8_t* front,
                   vo_counters* counters) {
    const Scene* s = (const Scene*)p;
    Counters c;
    for (uint64_t i = 0; i < n; ++i) {
        Ray ray(v3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
        HitRecord h;
        size_t sf = 0;
        if (s->scene_hit(ray, h, sf, c, mode)) {
            surface[i] = (uint32_t)sf;
            prim[i] = h.prim;
            t[i] = h.t;
            if (normal) { normal[3 * i] = h.normal.x; normal[3 * i + 1] = h.normal.y; normal[3 * i + 2] = h.normal.z; }
            if (uv) { uv[2 * i] = h.uv.x; uv[2 * i + 1] = h.uv.y; }
            if (front) front[i] = h.front_face ? 1 : 0;
        } else {
            surface[i] = 0xFFFFFFFFu;
            prim[i] = 0xFFFFFFFFu;
            t[i] = INF_F;
            if (normal) { normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0.0f; }
            if (uv) { uv[2 * i] = uv[2 * i + 1] = 0.0f; }
            if (front) front[i] = 0;
        }
    }
    if (counters) { counters->box_tests = c.box_tests; counters->tri_tests = c.tri_tests; counters->segments = c.segments; }
}

// The camera ray of global sample `sample` of every pixel (jitter + DOF drawn from the sample's own
// stream) and its closest hit: the primary-ray gate.
void vo_trace_primary(void* p, uint32_t W, uint32_t H, const vo_settings* vs, uint32_t sample,
                      float* origins, float* dirs, uint32_t* surface, uint32_t* prim, float* t,
                      vo_counters* counters) {
    const Scene* s = (const Scene*)p;
    const RenderSettings st = to_settings(vs);
    Counters c;
    for (uint32_t index = 0; index < W * H; ++index) {
        F x, y, d;
        pixel_to_ndc(index, W, H, st.pixel_mapping, x, y, d);
        Rng rng(st.seed, index, sample);
        const Ray ray = camera_sample_ray(*s, x, y, d, rng);
        if (origins) { origins[3 * index] = ray.origin.x; origins[3 * index + 1] = ray.origin.y; origins[3 * index + 2] = ray.origin.z; }
        if (dirs) { dirs[3 * index] = ray.direction.x; dirs[3 * index + 1] = ray.direction.y; dirs[3 * index + 2] = ray.direction.z; }
        HitRecord h;
        size_t sf = 0;
        if (s->scene_hit(ray, h, sf, c, st.traverse_mode)) {
            surface[index] = (uint32_t)sf;
            prim[index] = h.prim;
            t[index] = h.t;
        } else {
            surface[index] = 0xFFFFFFFFu;
            prim[index] = 0xFFFFFFFFu;
            t[index] = INF_F;
        }
    }
    if (counters) { counters->box_tests = c.box_tests; counters->tri_tests = c.tri_tests; counters->segments = c.segments; }
}

// Radiance of single camera samples (debug / per-sample parity): out[3*i..] for (pixel[i], sample[i]).
void vo_sample_radiance(void* p, uint32_t W, uint32_t H, const vo_settings* vs, uint64_t n,
                        const uint32_t* pixel, const uint32_t* sample, float* out) {
    const Scene* s = (const Scene*)p;
    const RenderSettings st = to_settings(vs);
    Counters c;
    for (uint64_t i = 0; i < n; ++i) {
        F x, y, d;
        pixel_to_ndc(pixel[i], W, H, st.pixel_mapping, x, y, d);
        Rng rng(st.seed, pixel[i], sample[i]);
        const Ray ray = camera_sample_ray(*s, x, y, d, rng);
        const Color col = trace_ray_internal(*s, st, ray, 0, rng, c);
        out[3 * i] = col.x; out[3 * i + 1] = col.y; out[3 * i + 2] = col.z;
    }
}

// iterative_render (render/iterative.rs:11-55) over pixel range [pixel_begin, pixel_end) of a W x H
// target: for each pixel, `samples` camera samples with global indices sample_offset.., summed,
// scaled by 1/total_samples and added into accum (RGBA f32, alpha += 1 per call). n_threads workers
// pull 256-pixel chunks from a shared counter (stand-in for rayon's par_chunks_exact_mut(4)).
void vo_render(void* p, uint32_t W, uint32_t H, const vo_settings* vs, uint32_t sample_offset,
               uint32_t samples, uint32_t pixel_begin, uint32_t pixel_end, float* accum,
               int32_t n_threads, vo_counters* counters) {
    const Scene* s = (const Scene*)p;
    const RenderSettings st = to_settings(vs);
    if (n_threads < 1) n_threads = 1;
    std::atomic<uint32_t> next(pixel_begin);
    std::atomic<uint64_t> box(0), tri(0), seg(0);
    const uint32_t CHUNK = 256;
    auto worker = [&]() {
        Counters c;
        while (true) {
            const uint32_t begin = next.fetch_add(CHUNK);
            if (begin >= pixel_end) break;
            const uint32_t end = std::min(begin + CHUNK, pixel_end);
            for (uint32_t index = begin; index < end; ++index) {
                F x, y, d;
                pixel_to_ndc(index, W, H, st.pixel_mapping, x, y, d);
                Color color = BLACK;
                for (uint32_t k = 0; k < samples; ++k) {
                    Rng rng(st.seed, index, sample_offset + k);
                    const Ray ray = camera_sample_ray(*s, x, y, d, rng);
                    color = color + trace_ray_internal(*s, st, ray, 0, rng, c);
                }
                color = color * (1.0f / (F)st.total_samples);
                float* px = accum + 4 * (size_t)index;
                px[0] += color.x;
                px[1] += color.y;
                px[2] += color.z;
                px[3] += 1.0f;  // Color::a() == 1.0, color.rs:54
            }
        }
        box += c.box_tests;
        tri += c.tri_tests;
        seg += c.segments;
    };
    std::vector<std::thread> threads;
    for (int i = 1; i < n_threads; ++i) threads.emplace_back(worker);
    worker();
    for (auto& t : threads) t.join();
    if (counters) { counters->box_tests = box; counters->tri_tests = tri; counters->segments = seg; }
}

// post_process.glsl:23-49 + tonemapping.glsl:2-40 restated per pixel. GLSL mat3(a..i) is column-major.
static inline V3 mat3_mul(const F m[9], V3 v) {
    return V3{m[0] * v.x + m[3] * v.y + m[6] * v.z, m[1] * v.x + m[4] * v.y + m[7] * v.z,
              m[2] * v.x + m[5] * v.y + m[8] * v.z};
}
void vo_resolve(const float* accum, uint32_t W, uint32_t H, float scale, float gamma, float exposure,
                int32_t tonemap, float* out) {
    static const F ACES_IN[9] = {0.59719f, 0.076f, 0.0284f, 0.35458f, 0.90834f, 0.13383f, 0.04823f, 0.01566f, 0.83777f};
    static const F ACES_OUT[9] = {1.60475f, -0.10208f, -0.00327f, -0.53108f, 1.10813f, -0.07276f, -0.07367f, -0.00605f, 1.07602f};
    const F e = std::pow(2.0f, exposure);
    const F inv_gamma = 1.0f / gamma;
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        V3 c = v3(accum[4 * i] * scale, accum[4 * i + 1] * scale, accum[4 * i + 2] * scale) * e;
        switch (tonemap) {
            case 1: {  // ACES (Hill fit)
                c = mat3_mul(ACES_IN, c);
                V3 r;
                r.x = (c.x * (c.x + 0.0245786f) - 0.000090537f) / (c.x * (0.983729f * c.x + 0.432951f) + 0.238081f);
                r.y = (c.y * (c.y + 0.0245786f) - 0.000090537f) / (c.y * (0.983729f * c.y + 0.432951f) + 0.238081f);
                r.z = (c.z * (c.z + 0.0245786f) - 0.000090537f) / (c.z * (0.983729f * c.z + 0.432951f) + 0.238081f);
                c = mat3_mul(ACES_OUT, r);
                break;
            }
            case 2: {  // Reinhard, white = 2
                const F white = 2.0f;
                const F luma = (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f;
                const F tm = luma * (1.0f + luma / (white * white)) / (1.0f + luma);
                c = c * (tm / luma);
                break;
            }
            case 3: {  // Filmic (Hejl / Burgess-Dawson)
                c = v3(rmax(0.0f, c.x - 0.004f), rmax(0.0f, c.y - 0.004f), rmax(0.0f, c.z - 0.004f));
                c.x = (c.x * (6.2f * c.x + 0.5f)) / (c.x * (6.2f * c.x + 1.7f) + 0.06f);
                c.y = (c.y * (6.2f * c.y + 0.5f)) / (c.y * (6.2f * c.y + 1.7f) + 0.06f);
                c.z = (c.z * (6.2f * c.z + 0.5f)) / (c.z * (6.2f * c.z + 1.7f) + 0.06f);
                break;
            }
            case 4: {  // Uncharted 2
                const F A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, Fc = 0.30f, Wp = 11.2f;
                c = c * 2.0f;
                c.x = ((c.x * (A * c.x + C * B) + D * E) / (c.x * (A * c.x + B) + D * Fc)) - E / Fc;
                c.y = ((c.y * (A * c.y + C * B) + D * E) / (c.y * (A * c.y + B) + D * Fc)) - E / Fc;
                c.z = ((c.z * (A * c.z + C * B) + D * E) / (c.z * (A * c.z + B) + D * Fc)) - E / Fc;
                const F white = ((Wp * (A * Wp + C * B) + D * E) / (Wp * (A * Wp + B) + D * Fc)) - E / Fc;
                c = c / white;
                break;
            }
            default:
                break;
        }
        out[4 * i] = std::pow(c.x, inv_gamma);
        out[4 * i + 1] = std::pow(c.y, inv_gamma);
        out[4 * i + 2] = std::pow(c.z, inv_gamma);
        out[4 * i + 3] = 1.0f;
    }
}

// Known-answer hooks for the generator and samplers.
void vo_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void vo_rng_draws(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out) {
    Rng rng(seed, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) out[i] = rng.next_u32();
}
void vo_unit_sphere(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out) {
    Rng rng(seed, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) {
        const V3 v = rng.unit_sphere();
        out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
}
void vo_unit_disc(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out) {
    Rng rng(seed, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) {
        const V2 v = rng.unit_disc();
        out[2 * i] = v.x; out[2 * i + 1] = v.y;
    }
}
void vo_texture_sample(void* p, uint32_t tex, uint64_t n, const float* uv, float* out) {
    const Scene* s = (const Scene*)p;
    for (uint64_t i = 0; i < n; ++i) {
        const Color c = s->textures[tex].sample(uv[2 * i], uv[2 * i + 1]);
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
void vo_environment_sample(void* p, uint64_t n, const float* dirs, float* out) {
    const Scene* s = (const Scene*)p;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r;
        r.origin = v3(0, 0, 0);
        r.direction = v3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        const Color c = s->environment ? s->environment->sample(r) : BLACK;
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}
float vo_schlick(float cosine, float idx) { return Dielectric::reflectance(cosine, idx); }

}  // extern "C"
