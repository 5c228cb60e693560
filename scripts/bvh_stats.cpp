// scripts/bvh_stats.cpp — offline BVH quality meter (host only, not part of the library).
// Flattens a scene with the library's own builder (csrc/scene_build.cpp) and walks the resulting quantised BVH2
// on the CPU the way k_trace does (near child first, far child deferred, closest hit shrinks the interval),
// for camera rays and three generations of diffuse bounce rays. Prints node visits and triangle tests per ray
// and the tree's SAH cost, so that builder changes can be compared without a GPU.
//
//   g++ -O2 -std=c++17 -pthread -ffp-contract=off -Ivoidray_b200/csrc -x c++ voidray_b200/csrc/scene_build.cpp
//       scripts/bvh_stats.cpp -o /tmp/bvh_stats   (one command line)
//   /tmp/bvh_stats assets/mossy_ground.obj [copies_x copies_z]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "scene_build.h"

using namespace vr;

struct Ray {
    float o[3], d[3];
};
struct Hit {
    float t;
    int tri;
};

static uint32_t g_state = 12345u;
static float rnd() {
    g_state = g_state * 1664525u + 1013904223u;
    return (float)(g_state >> 8) * (1.0f / 16777216.0f);
}

struct Tree {
    const FlatScene& f;
    float cell[3];
    explicit Tree(const FlatScene& fs) : f(fs) {
        for (int a = 0; a < 3; ++a) cell[a] = fs.grid_extent[a] / 32768.0f;
    }
    void child_box(size_t node, int c, float lo[3], float hi[3]) const {
        const Quad* q = &f.nodes[node * NODE_QUADS];
        uint32_t w[6];
        const float src[6] = {q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y};
        std::memcpy(w, src, 24);
        for (int a = 0; a < 3; ++a) {
            const uint32_t p = w[3 * c + a];
            lo[a] = f.grid_min[a] + (float)(p & 0x7FFF) * cell[a];
            hi[a] = f.grid_min[a] + (float)((p >> 16) & 0x7FFF) * cell[a];
        }
    }
    int child_code(size_t node, int c) const {
        const Quad* q = &f.nodes[node * NODE_QUADS];
        float v = c == 0 ? q[1].z : q[1].w;
        int code;
        std::memcpy(&code, &v, 4);
        return code;
    }
    bool slab(const float lo[3], const float hi[3], const Ray& r, const float inv[3], float tmax, float& tn) const {
        float t0 = 0.0f, t1 = tmax;
        for (int a = 0; a < 3; ++a) {
            float ta = (lo[a] - r.o[a]) * inv[a], tb = (hi[a] - r.o[a]) * inv[a];
            if (ta > tb) std::swap(ta, tb);
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
        }
        tn = t0;
        return t0 <= t1 * 1.0000005f;
    }
    bool tri_hit(int i, const Ray& r, float& t) const {
        const Quad* q = &f.tri_isect[(size_t)i * TRI_ISECT_QUADS];
        const float v0[3] = {q[0].x, q[0].y, q[0].z}, e1[3] = {q[1].x, q[1].y, q[1].z}, e2[3] = {q[2].x, q[2].y, q[2].z};
        const float h[3] = {r.d[1] * e2[2] - r.d[2] * e2[1], r.d[2] * e2[0] - r.d[0] * e2[2], r.d[0] * e2[1] - r.d[1] * e2[0]};
        const float a = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
        if (a > -1e-5f && a < 1e-5f) return false;
        const float fi = 1.0f / a;
        const float s[3] = {r.o[0] - v0[0], r.o[1] - v0[1], r.o[2] - v0[2]};
        const float u = fi * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
        if (u < 0.0f || u > 1.0f) return false;
        const float qq[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
        const float v = fi * (r.d[0] * qq[0] + r.d[1] * qq[1] + r.d[2] * qq[2]);
        if (v < 0.0f || u + v > 1.0f) return false;
        t = fi * (e2[0] * qq[0] + e2[1] * qq[1] + e2[2] * qq[2]);
        return t > 1e-5f;
    }
    Hit trace(const Ray& r, uint64_t& n_nodes, uint64_t& n_tris, uint32_t& max_sp) const {
        Hit best{INFINITY, -1};
        float inv[3];
        for (int a = 0; a < 3; ++a) inv[a] = 1.0f / r.d[a];
        int stack[64];
        int sp = 0;
        int cur = 0;
        for (;;) {
            if (cur >= 0) {
                ++n_nodes;
                float lo[3], hi[3], ta = 0, tb = 0;
                child_box(cur, 0, lo, hi);
                const bool ha = slab(lo, hi, r, inv, best.t, ta);
                child_box(cur, 1, lo, hi);
                const bool hb = slab(lo, hi, r, inv, best.t, tb);
                const int ca = child_code(cur, 0), cb = child_code(cur, 1);
                const bool b_first = hb && (!ha || tb < ta);
                const int near_c = b_first ? cb : ca, far_c = b_first ? ca : cb;
                if (ha && hb) {
                    stack[sp++] = far_c;
                    if ((uint32_t)sp > max_sp) max_sp = sp;
                }
                if (ha || hb) {
                    cur = near_c;
                    continue;
                }
            } else {
                const uint32_t code = ~(uint32_t)cur;
                const uint32_t first = code >> 3, count = code & 7;
                for (uint32_t k = 0; k < count; ++k) {
                    ++n_tris;
                    float t;
                    if (tri_hit((int)(first + k), r, t) && t < best.t) best = Hit{t, (int)(first + k)};
                }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
        return best;
    }
    double sah_cost(float node_cost) const {
        // sum over inner nodes of area(child) / area(root) * (node_cost or n_tris of a leaf child)
        float lo[3], hi[3];
        auto area = [&](const float* l, const float* h) {
            const float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
            return dx < 0 || dy < 0 || dz < 0 ? 0.0 : (double)(dx * dy + dy * dz + dz * dx);
        };
        float rl[3] = {INFINITY, INFINITY, INFINITY}, rh[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int c = 0; c < 2; ++c) {
            child_box(0, c, lo, hi);
            if (lo[0] > hi[0]) continue;
            for (int a = 0; a < 3; ++a) { rl[a] = std::min(rl[a], lo[a]); rh[a] = std::max(rh[a], hi[a]); }
        }
        const double root = area(rl, rh);
        double cost = node_cost;
        std::vector<int> st{0};
        while (!st.empty()) {
            const int n = st.back();
            st.pop_back();
            for (int c = 0; c < 2; ++c) {
                child_box(n, c, lo, hi);
                const int code = child_code(n, c);
                const double p = area(lo, hi) / root;
                if (code >= 0) { cost += p * node_cost; st.push_back(code); }
                else cost += p * (double)((~(uint32_t)code) & 7);
            }
        }
        return cost;
    }
};


// ---- L1 model (BVH_STATS_CACHE=<KB>): how the node order decides what an SM's L1 holds -----------------------------
// 1024 rays in flight (32 warps x 32 lanes) take one traversal step each per round, a finished ray is replaced by the
// next one of the list; every node visit touches one 32-byte sector, every triangle test two. The cache is a fully
// associative LRU over 128-byte lines with per-sector valid bits (a line is allocated on its first sector).
// Node orders for the L1 model. perm[new index] = old index; node 0 (a copy of the root, node 1) stays.
static void reorder_nodes(FlatScene& f, const std::vector<uint32_t>& perm) {
    const size_t n = f.nodes.size() / NODE_QUADS;
    std::vector<uint32_t> where(n);
    for (size_t i = 0; i < n; ++i) where[perm[i]] = (uint32_t)i;
    RawVector<Quad> out(f.nodes.size());
    for (size_t i = 0; i < n; ++i) {
        Quad a = f.nodes[(size_t)perm[i] * NODE_QUADS], b = f.nodes[(size_t)perm[i] * NODE_QUADS + 1];
        int32_t c0, c1;
        std::memcpy(&c0, &b.z, 4);
        std::memcpy(&c1, &b.w, 4);
        if (c0 >= 0) { c0 = (int32_t)where[c0]; std::memcpy(&b.z, &c0, 4); }
        if (c1 >= 0) { c1 = (int32_t)where[c1]; std::memcpy(&b.w, &c1, 4); }
        out[i * NODE_QUADS] = a;
        out[i * NODE_QUADS + 1] = b;
    }
    f.nodes.swap(out);
}
// area: nodes sorted by the surface area of their own box (the chance that a random ray visits them), largest first;
// bfs<k>: the top k levels breadth-first, the subtrees below them depth-first as before
static std::vector<uint32_t> node_order(const FlatScene& f, const std::string& kind) {
    const size_t n = f.nodes.size() / NODE_QUADS;
    Tree tree(f);
    std::vector<uint32_t> perm;
    perm.push_back(0);
    if (n < 2) return perm;
    std::vector<double> area(n, 0.0);
    std::vector<uint32_t> depth(n, 0);
    {
        // node 1 is the root (node 0 its copy); area of a node = area of the union of its two child boxes
        std::vector<uint32_t> st{1};
        depth[1] = 0;
        while (!st.empty()) {
            const uint32_t i = st.back();
            st.pop_back();
            float lo[3], hi[3], l2[3], h2[3];
            tree.child_box(i, 0, lo, hi);
            tree.child_box(i, 1, l2, h2);
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], l2[a]); hi[a] = std::max(hi[a], h2[a]); }
            const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
            area[i] = dx * dy + dy * dz + dz * dx;
            for (int c = 0; c < 2; ++c) {
                const int code = tree.child_code(i, c);
                if (code >= 0) { depth[code] = depth[i] + 1; st.push_back((uint32_t)code); }
            }
        }
    }
    std::vector<uint32_t> rest;
    for (uint32_t i = 1; i < n; ++i) rest.push_back(i);
    if (kind == "area") {
        std::stable_sort(rest.begin(), rest.end(), [&](uint32_t a, uint32_t b) { return area[a] > area[b]; });
    } else if (kind.rfind("bfs", 0) == 0) {
        const uint32_t k = (uint32_t)atoi(kind.c_str() + 3);
        std::stable_sort(rest.begin(), rest.end(), [&](uint32_t a, uint32_t b) {
            const uint32_t da = std::min(depth[a], k), db = std::min(depth[b], k);
            return da < db;  // levels < k breadth-first (stable: left to right), everything deeper keeps the DFS order
        });
    }
    perm.insert(perm.end(), rest.begin(), rest.end());
    return perm;
}

// ---- kernel-exact walks: the slab arithmetic of k_trace (PRMT-decoded plane, one FMA per plane, per-axis error
// bounds, best-t culling, rank tie rule) restated op for op, over the shipped BVH2 records and over the 4-wide nodes of
// collapse_bvh4 (experiment -DVR_BVH4). Both must report the brute-force closest hit bit for bit; the step counts are
// what the two traversals cost.
struct KHit {
    float t;
    int tri;
    uint32_t rank;
};
struct KernelWalk {
    const FlatScene& f;
    const RawVector<Quad>& wide;
    float a[3], bn[3], bf[3], o[3], d[3];
    bool pos[3];
    KHit best;
    KernelWalk(const FlatScene& fs, const RawVector<Quad>& w) : f(fs), wide(w) {}
    static uint32_t bits(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
    static float plane(uint32_t half, float a_, float b_) {
        const uint32_t u = 0x3F000000u | ((half & 0xFFFFu) << 8);
        float x; std::memcpy(&x, &u, 4);
        return std::fmaf(x, a_, b_);
    }
    void begin(const Ray& r) {
        for (int k = 0; k < 3; ++k) {
            o[k] = r.o[k]; d[k] = r.d[k];
            const float tiny = 1e-20f;
            const float id = 1.0f / (std::fabs(d[k]) > tiny ? d[k] : std::copysign(tiny, d[k]));
            a[k] = f.grid_extent[k] * id;
            const float g = (f.grid_min[k] - o[k]) * id;
            const float b = g - a[k];
            const float err = 2.4e-7f * (std::fabs(g) + std::fabs(a[k])) + 1e-30f;
            bn[k] = b - err; bf[k] = b + err; pos[k] = id >= 0.0f;
        }
        best = KHit{INFINITY, -1, 0};
    }
    bool slab(const uint32_t w[3], float& tn) const {
        float n = 0.0f, fr = best.t;
        float nn[3], ff[3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t lo = w[k] & 0xFFFFu, hi = w[k] >> 16;
            nn[k] = plane(pos[k] ? lo : hi, a[k], bn[k]);
            ff[k] = plane(pos[k] ? hi : lo, a[k], bf[k]);
        }
        n = std::fmax(std::fmax(nn[0], nn[1]), std::fmax(nn[2], 0.0f));
        fr = std::fmin(std::fmin(ff[0], ff[1]), std::fmin(ff[2], best.t));
        tn = n;
        return n <= fr * 1.0000005f;
    }
    void leaf(int code_neg, uint64_t& n_tris) {
        const uint32_t code = ~(uint32_t)code_neg;
        const uint32_t first = code >> 3, count = code & 7;
        for (uint32_t k = 0; k < count; ++k) {
            ++n_tris;
            const int i = (int)(first + k);
            const Quad* q = &f.tri_isect[(size_t)i * TRI_ISECT_QUADS];
            const float v0[3] = {q[0].x, q[0].y, q[0].z}, e1[3] = {q[1].x, q[1].y, q[1].z}, e2[3] = {q[2].x, q[2].y, q[2].z};
            const float h[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
            const float det = e1[0] * h[0] + e1[1] * h[1] + e1[2] * h[2];
            if (det > -1e-5f && det < 1e-5f) continue;
            const float fi = 1.0f / det;
            const float s[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
            const float u = fi * (s[0] * h[0] + s[1] * h[1] + s[2] * h[2]);
            if (u < 0.0f || u > 1.0f) continue;
            const float qq[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
            const float v = fi * (d[0] * qq[0] + d[1] * qq[1] + d[2] * qq[2]);
            if (v < 0.0f || u + v > 1.0f) continue;
            const float t = fi * (e2[0] * qq[0] + e2[1] * qq[1] + e2[2] * qq[2]);
            if (!(t > 1e-5f)) continue;
            const uint32_t rank = bits(q[0].w);
            if (t < best.t || (t == best.t && rank > best.rank)) best = KHit{t, i, rank};
        }
    }
    // k_trace over the BVH2 records (trav_node / trav_leaf_step of kernels.cu)
    KHit trace2(const Ray& r, uint64_t& n_nodes, uint64_t& n_tris, uint32_t& max_sp) {
        begin(r);
        int stack[128], sp = 0, cur = f.n_tris ? 0 : 0x7FFFFFFF;
        while (cur != 0x7FFFFFFF) {
            if (cur >= 0) {
                ++n_nodes;
                const Quad* q = &f.nodes[(size_t)cur * NODE_QUADS];
                const uint32_t w[8] = {bits(q[0].x), bits(q[0].y), bits(q[0].z), bits(q[0].w), bits(q[1].x), bits(q[1].y), bits(q[1].z), bits(q[1].w)};
                float ta, tb;
                const bool ha = slab(w, ta), hb = slab(w + 3, tb);
                const int ca = (int)w[6], cb = (int)w[7];
                const bool b_first = hb && (!ha || tb < ta);
                const int near_c = b_first ? cb : ca, far_c = b_first ? ca : cb;
                if (ha && hb) { stack[sp++] = far_c; max_sp = std::max(max_sp, (uint32_t)sp); }
                cur = (ha || hb) ? near_c : (sp ? stack[--sp] : 0x7FFFFFFF);
            } else {
                leaf(cur, n_tris);
                cur = sp ? stack[--sp] : 0x7FFFFFFF;
            }
        }
        return best;
    }
    // the -DVR_BVH4 trav_node: four slab tests per 64-byte node, hits sorted by entry distance, nearest first
    int sort_mode = 0;
    bool cull_on_pop = false;  // experiment: park the entry distance with the child and drop it at pop time if the closest hit got nearer
    KHit trace4(const Ray& r, uint64_t& n_nodes, uint64_t& n_tris, uint32_t& max_sp) {
        begin(r);
        int stack[WIDE_STACK_LIMIT + 4], sp = 0, cur = f.n_tris ? 0 : 0x7FFFFFFF;
        float kstack[WIDE_STACK_LIMIT + 4];
        auto pop = [&]() {
            while (sp) {
                --sp;
                if (!cull_on_pop || kstack[sp] <= best.t * 1.0000005f) return stack[sp];
            }
            return 0x7FFFFFFF;
        };
        while (cur != 0x7FFFFFFF) {
            if (cur >= 0) {
                ++n_nodes;
                const Quad* q = &wide[(size_t)cur * WIDE_NODE_QUADS];
                float key[4];
                int code[4];
                int h = 0;
                for (int p = 0; p < 2; ++p) {
                    const Quad* r2 = q + 2 * p;
                    const uint32_t w[8] = {bits(r2[0].x), bits(r2[0].y), bits(r2[0].z), bits(r2[0].w), bits(r2[1].x), bits(r2[1].y), bits(r2[1].z), bits(r2[1].w)};
                    for (int c = 0; c < 2; ++c) {
                        float tn;
                        const bool hit = slab(w + 3 * c, tn);
                        key[2 * p + c] = hit ? tn : INFINITY;
                        code[2 * p + c] = (int)w[6 + c];
                        h += hit ? 1 : 0;
                    }
                }
                auto cswap = [&](int i, int j) {
                    if (key[j] < key[i]) { std::swap(key[i], key[j]); std::swap(code[i], code[j]); }
                };
                if (sort_mode == 0) {  // the kernel's 5-comparator network
                    cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);
                } else {  // experiment: only the nearest child is found, the others keep their slot order (misses last)
                    int m = 0;
                    for (int i = 1; i < 4; ++i) if (key[i] < key[m]) m = i;
                    std::swap(key[0], key[m]); std::swap(code[0], code[m]);
                    // stable partition of slots 1..3: hits first
                    for (int pass = 0; pass < 2; ++pass)
                        for (int i = 1; i < 3; ++i)
                            if (!(key[i] < INFINITY) && key[i + 1] < INFINITY) { std::swap(key[i], key[i + 1]); std::swap(code[i], code[i + 1]); }
                }
                if (h > 3) { kstack[sp] = key[3]; stack[sp++] = code[3]; }
                if (h > 2) { kstack[sp] = key[2]; stack[sp++] = code[2]; }
                if (h > 1) { kstack[sp] = key[1]; stack[sp++] = code[1]; }
                max_sp = std::max(max_sp, (uint32_t)sp);
                cur = h > 0 ? code[0] : pop();
            } else {
                leaf(cur, n_tris);
                cur = pop();
            }
        }
        return best;
    }
    KHit brute(const Ray& r) {
        begin(r);
        uint64_t n = 0;
        for (uint32_t i = 0; i < f.n_tris; ++i) leaf(~(int)((i << 3) | 1u), n);
        return best;
    }
};

struct Walker {
    const Tree* tree;
    Ray r;
    float inv[3];
    Hit best;
    int stack[64];
    int sp, cur, leaf_k;
    bool done;
    void start(const Tree* t, const Ray& ray) {
        tree = t; r = ray; best = Hit{INFINITY, -1}; sp = 0; cur = 0; leaf_k = 0; done = false;
        for (int a = 0; a < 3; ++a) inv[a] = 1.0f / r.d[a];
    }
    // one step; returns the byte address touched (nodes from 0, triangle records from 1 << 40) and its length
    void step(uint64_t& addr, uint32_t& len) {
        if (cur >= 0) {
            addr = (uint64_t)cur * 32; len = 32;
            float lo[3], hi[3], ta = 0, tb = 0;
            tree->child_box(cur, 0, lo, hi);
            const bool ha = tree->slab(lo, hi, r, inv, best.t, ta);
            tree->child_box(cur, 1, lo, hi);
            const bool hb = tree->slab(lo, hi, r, inv, best.t, tb);
            const int ca = tree->child_code(cur, 0), cb = tree->child_code(cur, 1);
            const bool b_first = hb && (!ha || tb < ta);
            const int near_c = b_first ? cb : ca, far_c = b_first ? ca : cb;
            if (ha && hb) stack[sp++] = far_c;
            if (ha || hb) { cur = near_c; leaf_k = 0; return; }
        } else {
            const uint32_t code = ~(uint32_t)cur;
            const uint32_t first = code >> 3, count = code & 7;
            if ((uint32_t)leaf_k < count) {
                addr = ((uint64_t)1 << 40) + (uint64_t)(first + leaf_k) * 64; len = 64;
                float t;
                if (tree->tri_hit((int)(first + leaf_k), r, t) && t < best.t) best = Hit{t, (int)(first + leaf_k)};
                if ((uint32_t)++leaf_k < count) return;
            } else { addr = 0; len = 0; }
        }
        if (sp == 0) { done = true; return; }
        cur = stack[--sp]; leaf_k = 0;
    }
};

struct SectorCache {
    struct Line { uint64_t tag; uint32_t valid; uint64_t stamp; };
    std::vector<Line> lines;
    std::vector<std::pair<uint64_t, uint32_t>> index;  // open-addressing map tag -> slot (slot + 1, 0 = empty)
    uint64_t clock = 0, hits = 0, misses = 0;
    explicit SectorCache(size_t kb) : lines(kb * 1024 / 128, Line{~0ull, 0, 0}) {}
    void touch(uint64_t addr, uint32_t len) {
        for (uint64_t a = addr & ~31ull; a < addr + len; a += 32) {
            const uint64_t tag = a >> 7;
            const uint32_t bit = 1u << ((a >> 5) & 3);
            ++clock;
            Line* found = nullptr;
            Line* lru = &lines[0];
            for (Line& l : lines) {  // small (a few hundred lines): a linear scan is fine for an offline meter
                if (l.tag == tag) { found = &l; break; }
                if (l.stamp < lru->stamp) lru = &l;
            }
            if (found) {
                found->stamp = clock;
                if (found->valid & bit) ++hits; else { ++misses; found->valid |= bit; }
            } else {
                ++misses;
                *lru = Line{tag, bit, clock};
            }
        }
    }
};

static void cache_model(const Tree& tree, const std::vector<Ray>& rays, size_t kb, const char* what) {
    SectorCache cache(kb);
    const size_t IN_FLIGHT = 1024;
    std::vector<Walker> w(std::min(IN_FLIGHT, rays.size()));
    size_t next = 0;
    for (Walker& x : w) x.start(&tree, rays[next++]);
    size_t live = w.size();
    while (live) {
        for (Walker& x : w) {
            if (x.done) continue;
            uint64_t addr = 0; uint32_t len = 0;
            x.step(addr, len);
            if (len) cache.touch(addr, len);
            if (x.done) {
                if (next < rays.size()) x.start(&tree, rays[next++]);
                else --live;
            }
        }
    }
    std::printf("  L1 model %zu KB, %s: %.1f %% sector hit rate (%llu sector accesses)\n", kb, what,
                100.0 * cache.hits / std::max<uint64_t>(1, cache.hits + cache.misses), (unsigned long long)(cache.hits + cache.misses));
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const int nx = argc > 2 ? atoi(argv[2]) : 1, nz = argc > 3 ? atoi(argv[3]) : 1;
    HostMesh base;
    std::string err;
    if (!load_obj_file(argv[1], base, err)) { std::printf("%s\n", err.c_str()); return 1; }
    HostScene sc;
    HostMesh big;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < nz; ++j) {
            const float ang = 6.2831853f * (float)((i * 7919 + j * 104729) % 1000) / 1000.0f, c = cosf(ang), s = sinf(ang);
            const float ox = 1.5f * (float)i, oz = 1.5f * (float)j;
            const uint32_t v0 = (uint32_t)(big.pos.size() / 3);
            for (uint32_t v = 0; v < base.n_vertices; ++v) {
                const float x = base.pos[3 * v], y = base.pos[3 * v + 1], z = base.pos[3 * v + 2];
                if (nx * nz == 1) { big.pos.push_back(x); big.pos.push_back(y); big.pos.push_back(z); }
                else { big.pos.push_back(c * x + s * z + ox); big.pos.push_back(y); big.pos.push_back(-s * x + c * z + oz); }
                big.uv.push_back(0); big.uv.push_back(0);
                big.nrm.push_back(0); big.nrm.push_back(1); big.nrm.push_back(0);
            }
            for (uint32_t k : base.idx) big.idx.push_back(v0 + k);
        }
    big.n_vertices = (uint32_t)(big.pos.size() / 3);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t v = 0; v < big.n_vertices; ++v)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], big.pos[3 * v + a]); hi[a] = std::max(hi[a], big.pos[3 * v + a]); }
    const float lo_all[3] = {lo[0], lo[1], lo[2]}, hi_all[3] = {hi[0], hi[1], hi[2]};
    sc.meshes.push_back(std::move(big));
    HostSurface sf; sf.kind = 0; sf.mesh = 0; sc.surfaces.push_back(sf);
    MaterialRec m{}; m.albedo_tex = -1; m.normal_tex = -1; sc.materials.push_back(m);
    sc.objects.push_back(HostObject{0, 0});
    const float ce[3] = {0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2])};
    const float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
    const float eye[3] = {ce[0] + 0.2f * ext, ce[1] + 0.9f * ext, ce[2] - 1.6f * ext}, up[3] = {0, 1, 0};
    std::memcpy(sc.camera.eye, eye, 12);
    camera_look_at(eye, ce, up, sc.camera.direction, sc.camera.up);
    sc.camera.fov = 0.6f; sc.camera.has_dof = 0;
    FlatScene flat;
    const auto t0 = std::chrono::steady_clock::now();
    if (!flatten_scene(sc, flat, err)) { std::printf("%s\n", err.c_str()); return 1; }
    const double build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (const char* e = std::getenv("BVH_STATS_REPEAT")) {
        // warm repeat timing (a commit is repeated every bench step): minimum and median of N more flattens
        std::vector<double> ms;
        for (int k = 0, n = atoi(e); k < n; ++k) {
            FlatScene again;
            const auto r0 = std::chrono::steady_clock::now();
            if (!flatten_scene(sc, again, err)) return 1;
            ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - r0).count());
        }
        std::sort(ms.begin(), ms.end());
        if (!ms.empty()) std::printf("flatten x%zu: min %.3f ms, median %.3f ms\n", ms.size(), ms[0], ms[ms.size() / 2]);
    }
    {
        // FNV-1a of what the device receives: the same scene must flatten to the same bytes on every run
        auto fnv = [](const void* ptr, size_t n, uint64_t d) {
            const unsigned char* b = (const unsigned char*)ptr;
            for (size_t i = 0; i < n; ++i) d = (d ^ b[i]) * 1099511628211ull;
            return d;
        };
        uint64_t d = 1469598103934665603ull;
        d = fnv(flat.nodes.data(), flat.nodes.size() * sizeof(Quad), d);
        d = fnv(flat.tri_isect.data(), flat.tri_isect.size() * sizeof(Quad), d);
        d = fnv(flat.tri_shade.data(), flat.tri_shade.size() * sizeof(Quad), d);
        std::printf("flatten digest %016llx\n", (unsigned long long)d);
    }
    if (const char* e = std::getenv("BVH_STATS_ORDER")) reorder_nodes(flat, node_order(flat, e));  // experiment: see node_order
    Tree tree(flat);
    // camera rays on a 256 x 256 grid, then three generations of diffuse bounces (origin = hit point, direction = a
    // random unit vector flipped into the hemisphere facing back along the ray)
    std::vector<Ray> rays;
    const float* dir = sc.camera.direction;
    const float* cup = sc.camera.up;
    const float right[3] = {dir[1] * cup[2] - dir[2] * cup[1], dir[2] * cup[0] - dir[0] * cup[2], dir[0] * cup[1] - dir[1] * cup[0]};
    const float dd = 1.0f / std::tan(sc.camera.fov / 2.0f);
    const int G = 256;
    for (int y = 0; y < G; ++y)
        for (int x = 0; x < G; ++x) {
            const float px = ((float)x + 0.5f) / G * 2.0f - 1.0f, py = 1.0f - ((float)y + 0.5f) / G * 2.0f;
            Ray r;
            for (int a = 0; a < 3; ++a) { r.o[a] = eye[a]; r.d[a] = dd * dir[a] + px * right[a] + py * cup[a]; }
            rays.push_back(r);
        }
    std::printf("%s x%d: %u triangles, %zu nodes, depth %u, build %.1f ms, SAH cost %.2f\n", argv[1], nx * nz, flat.n_tris,
                flat.nodes.size() / NODE_QUADS, flat.bvh_depth, build_ms, tree.sah_cost(1.0f));
    uint64_t all_nodes = 0, all_tris = 0, all_rays = 0;
    std::vector<std::vector<Ray>> gen_rays;
    for (int gen = 0; gen < 4 && !rays.empty(); ++gen) {
        uint64_t n_nodes = 0, n_tris = 0, hits = 0;
        uint32_t max_sp = 0;
        std::vector<Ray> next;
        for (const Ray& r : rays) {
            const Hit h = tree.trace(r, n_nodes, n_tris, max_sp);
            if (h.tri < 0) continue;
            ++hits;
            Ray b;
            float v[3], len2;
            do {
                for (int a = 0; a < 3; ++a) v[a] = 2.0f * rnd() - 1.0f;
                len2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
            } while (len2 > 1.0f || len2 < 1e-6f);
            const float back = v[0] * r.d[0] + v[1] * r.d[1] + v[2] * r.d[2];
            for (int a = 0; a < 3; ++a) { b.o[a] = r.o[a] + r.d[a] * h.t; b.d[a] = back > 0 ? -v[a] : v[a]; }
            next.push_back(b);
        }
        std::printf("  generation %d: %zu rays, %.1f %% hit, %.2f nodes / ray, %.2f triangle tests / ray, max stack %u\n", gen,
                    rays.size(), 100.0 * hits / rays.size(), (double)n_nodes / rays.size(), (double)n_tris / rays.size(), max_sp);
        all_nodes += n_nodes; all_tris += n_tris; all_rays += rays.size();
        gen_rays.push_back(rays);
        if (const char* e = std::getenv("BVH_STATS_CACHE")) {
            if (gen <= 1) cache_model(tree, rays, (size_t)atoi(e), gen == 0 ? "camera rays" : "first-bounce rays");
        }
        rays.swap(next);
    }
    // the tree only culls: on a sample of the last generation's parents, the walk must find exactly the hit a loop
    // over every triangle finds
    {
        uint64_t dummy_n = 0, dummy_t = 0, mismatches = 0, checked = 0;
        uint32_t dummy_sp = 0;
        g_state = 777u;
        for (int k = 0; k < 3000; ++k) {
            Ray r;
            for (int a = 0; a < 3; ++a) { r.o[a] = lo_all[a] + rnd() * (hi_all[a] - lo_all[a]); r.d[a] = 2.0f * rnd() - 1.0f; }
            const Hit h = tree.trace(r, dummy_n, dummy_t, dummy_sp);
            Hit b{INFINITY, -1};
            for (uint32_t i = 0; i < flat.n_tris; ++i) {
                float t;
                if (tree.tri_hit((int)i, r, t) && t < b.t) b = Hit{t, (int)i};
            }
            ++checked;
            if (h.t != b.t) ++mismatches;
        }
        std::printf("  brute-force check: %llu of %llu random rays differ\n", (unsigned long long)mismatches, (unsigned long long)checked);
        if (mismatches) return 1;
    }
    if (const char* e = std::getenv("BVH_STATS_CACHE")) {
        // deep-bounce stand-in: random origins inside the scene box, random directions
        std::vector<Ray> random_rays;
        g_state = 4242u;
        for (int k = 0; k < 30000; ++k) {
            Ray r;
            for (int a = 0; a < 3; ++a) { r.o[a] = lo_all[a] + rnd() * (hi_all[a] - lo_all[a]); r.d[a] = 2.0f * rnd() - 1.0f; }
            random_rays.push_back(r);
        }
        cache_model(tree, random_rays, (size_t)atoi(e), "random rays");
    }
#ifndef VR_BVH4
    if (std::getenv("BVH_STATS_WIDE")) {
        // the 4-wide collapse (experiment -DVR_BVH4) next to the shipped BVH2, both walked with the kernel's own slab
        // arithmetic: identical hits, and what each costs in node fetches
        RawVector<Quad> wide;
        uint32_t wide_depth = 0, stack_bound = 0;
        const auto c0 = std::chrono::steady_clock::now();
        // BVH_STATS_WIDE=<n> with n > 1 collapses under a stack limit of n instead of the kernel's (checks that the
        // collapse narrows nodes to keep deep paths within the limit)
        const uint32_t limit = atoi(std::getenv("BVH_STATS_WIDE")) > 1 ? (uint32_t)atoi(std::getenv("BVH_STATS_WIDE")) : (uint32_t)WIDE_STACK_LIMIT;
        collapse_bvh4(flat.nodes, flat.grid_extent, flat.bvh_depth, limit, wide, &wide_depth, &stack_bound);
        if (stack_bound > std::max(limit, flat.bvh_depth)) { std::printf("stack bound %u exceeds the limit %u\n", stack_bound, limit); return 1; }
        const double collapse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
        size_t slots = 0;
        for (size_t i = 0; i < wide.size() / WIDE_NODE_QUADS; ++i)
            for (int p = 0; p < 2; ++p) {
                const Quad* q = &wide[i * WIDE_NODE_QUADS + 2 * p];
                slots += (KernelWalk::bits(q[1].z) != 0xFFFFFFFFu) + (KernelWalk::bits(q[1].w) != 0xFFFFFFFFu);
            }
        std::printf("  4-wide collapse: %zu nodes (%.2f children / node), depth %u, stack bound %u, %.2f ms\n", wide.size() / WIDE_NODE_QUADS,
                    (double)slots / std::max<size_t>(1, wide.size() / WIDE_NODE_QUADS), wide_depth, stack_bound, collapse_ms);
        KernelWalk kw(flat, wide);
        uint64_t differ = 0;
        for (size_t gen = 0; gen < gen_rays.size(); ++gen) {
            uint64_t n2 = 0, t2 = 0, n4 = 0, t4 = 0;
            uint32_t sp2 = 0, sp4 = 0;
            for (const Ray& r : gen_rays[gen]) {
                const KHit a = kw.trace2(r, n2, t2, sp2);
                const KHit b = kw.trace4(r, n4, t4, sp4);
                if (a.tri != b.tri || std::memcmp(&a.t, &b.t, 4) != 0) ++differ;
            }
            uint64_t n4c = 0, t4c = 0;
            if (std::getenv("BVH_STATS_NOSORT")) kw.sort_mode = 1; else kw.cull_on_pop = true;
            for (const Ray& r : gen_rays[gen]) {
                const KHit a = kw.trace2(r, n2, t2, sp2);
                const KHit b = kw.trace4(r, n4c, t4c, sp4);
                if (a.tri != b.tri || std::memcmp(&a.t, &b.t, 4) != 0) ++differ;
            }
            kw.cull_on_pop = false;
            kw.sort_mode = 0;
            n2 /= 2; t2 /= 2;
            const double nr = (double)std::max<size_t>(1, gen_rays[gen].size());
            std::printf("  generation %zu, kernel walk: BVH2 %.2f nodes + %.2f tris (stack %u) | BVH4 %.2f nodes + %.2f tris (stack %u) | "
                        "BVH4 + cull on pop %.2f nodes + %.2f tris\n", gen, n2 / nr, t2 / nr, sp2, n4 / nr, t4 / nr, sp4, n4c / nr, t4c / nr);
        }
        g_state = 777u;
        uint64_t brute_differ = 0;
        for (int k = 0; k < 3000; ++k) {
            Ray r;
            for (int a = 0; a < 3; ++a) { r.o[a] = lo_all[a] + rnd() * (hi_all[a] - lo_all[a]); r.d[a] = 2.0f * rnd() - 1.0f; }
            if (k % 7 == 0) r.d[k % 3] = 0.0f;  // axis-parallel rays
            uint64_t n = 0, t = 0;
            uint32_t sp = 0;
            const KHit a = kw.trace2(r, n, t, sp), b = kw.trace4(r, n, t, sp), c = kw.brute(r);
            if (a.tri != c.tri || std::memcmp(&a.t, &c.t, 4) != 0 || b.tri != c.tri || std::memcmp(&b.t, &c.t, 4) != 0) ++brute_differ;
        }
        std::printf("  kernel walks: %llu BVH2 / BVH4 differences, %llu of 3000 random rays differ from brute force\n",
                    (unsigned long long)differ, (unsigned long long)brute_differ);
        if (differ || brute_differ) return 1;
    }
#endif
    std::printf("  all: %.2f nodes / ray, %.2f triangle tests / ray, step estimate (nodes + 0.6 tris) %.2f\n", (double)all_nodes / all_rays,
                (double)all_tris / all_rays, ((double)all_nodes + 0.6 * all_tris) / all_rays);
    return 0;
}
