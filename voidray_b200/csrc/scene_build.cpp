// scene_build.cpp — flatten a HostScene into the packed device layout: world-space triangle records,
// a binned-SAH BVH2 whose nodes carry both child boxes, and the tie ranks that make the closest hit
// identical to the reference's (DESIGN.md §4 "Tie rule").
#include "scene_build.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <emmintrin.h>
#include <exception>
#include <system_error>
#include <thread>
#include <unordered_map>

namespace vr {
namespace {

struct V3 {
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
// cgmath 0.18 operation order: dot = (x*x' + y*y') + z*z'
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 normalize(V3 a) { return scale(a, 1.0f / std::sqrt(dot(a, a))); }
inline V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }

// f32::total_cmp key
inline int32_t total_key(float f) {
    int32_t i;
    std::memcpy(&i, &f, 4);
    i ^= (int32_t)(((uint32_t)(i >> 31)) >> 1);
    return i;
}
// f32::min / f32::max (NaN loses), inlined: std::fmin / std::fmax are libm calls at -O2
inline float fmin_(float a, float b) { return (a < b || b != b) ? a : b; }
inline float fmax_(float a, float b) { return (a > b || b != b) ? a : b; }
inline float bits_f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

}  // namespace

void camera_look_at(const float eye[3], const float center[3], const float up_in[3], float direction[3],
                    float up[3]) {
    const V3 dir = normalize(sub(ld3(center), ld3(eye)));
    const V3 u0 = ld3(up_in);
    const V3 u = normalize(sub(u0, scale(dir, dot(u0, dir))));
    direction[0] = dir.x; direction[1] = dir.y; direction[2] = dir.z;
    up[0] = u.x; up[1] = u.y; up[2] = u.z;
}

namespace {

inline size_t worker_count() { return std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32); }

// std::thread creation can fail (thread limits of a container): the caller then does the work itself
template <typename F>
bool try_spawn(std::vector<std::thread>& pool, F&& f) {
    try {
        pool.emplace_back(std::forward<F>(f));
        return true;
    } catch (const std::system_error&) {
        return false;
    }
}

// fn(begin, end) over [0, n) in contiguous chunks, one per hardware thread (inline when n is small)
template <typename Fn>
void parallel_for(size_t n, Fn fn) {
    const size_t T = n < 65536 ? 1 : worker_count();
    if (T == 1) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (n + T - 1) / T;
    for (size_t t = 0; t < T; ++t) {
        const size_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        if (!try_spawn(th, [=]() { fn(b, e); })) fn(b, e);
    }
    for (std::thread& x : th) x.join();
}

// fn(chunk index, begin, end): like parallel_for with the chunk number, for per-chunk partial results
template <typename Fn>
size_t parallel_chunks(size_t n, size_t T, Fn fn) {
    T = std::max<size_t>(1, std::min(T, n));
    const size_t chunk = (n + T - 1) / T;
    std::vector<std::thread> th;
    size_t used = 0;
    for (size_t t = 0; t < T; ++t) {
        const size_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        ++used;
        if (T == 1) fn(t, b, e);
        else if (!try_spawn(th, [=]() { fn(t, b, e); })) fn(t, b, e);
    }
    for (std::thread& x : th) x.join();
    return used;
}

// One item of the median-split recursion: its centroid travels with it, so every pass over a node streams.
struct CenRec {
    float c[3];
    uint32_t idx;
};
inline uint32_t sort_key(const CenRec& r, int axis) { return (uint32_t)total_key(r.c[axis]) ^ 0x80000000u; }

// Stable LSD radix sort (11 + 11 + 10 bits) of v[0, n) by the total_cmp key of c[axis]; tmp is scratch of the
// same length. T > 1: every pass histograms and scatters T contiguous chunks in parallel — chunk t's items of a
// digit go after chunk t-1's, so the result is the same stable order for any T.
void radix_sort_records(CenRec* v, CenRec* tmp, size_t n, int axis, size_t T) {
    T = std::max<size_t>(1, std::min<size_t>(T, n / 65536 + 1));
    std::vector<size_t> hist(T * 2048);
    CenRec* src = v;
    CenRec* dst = tmp;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass * 11;
        const uint32_t mask = pass == 2 ? 0x3FFu : 0x7FFu;
        std::fill(hist.begin(), hist.end(), (size_t)0);
        const size_t used = parallel_chunks(n, T, [&](size_t t, size_t b, size_t e) {
            size_t* h = &hist[t * 2048];
            for (size_t i = b; i < e; ++i) h[(sort_key(src[i], axis) >> shift) & mask]++;
        });
        size_t sum = 0;
        for (uint32_t d = 0; d <= mask; ++d)
            for (size_t t = 0; t < used; ++t) {
                const size_t c = hist[t * 2048 + d];
                hist[t * 2048 + d] = sum;
                sum += c;
            }
        parallel_chunks(n, T, [&](size_t t, size_t b, size_t e) {
            size_t* h = &hist[t * 2048];
            for (size_t i = b; i < e; ++i) dst[h[(sort_key(src[i], axis) >> shift) & mask]++] = src[i];
        });
        std::swap(src, dst);
    }
    // three passes: the result is in tmp
    parallel_chunks(n, T, [&](size_t, size_t b, size_t e) { std::memcpy(v + b, tmp + b, (e - b) * sizeof(CenRec)); });
}

inline void insertion_sort_records(CenRec* v, size_t n, int axis) {
    for (size_t i = 1; i < n; ++i) {
        const CenRec x = v[i];
        const int32_t kx = total_key(x.c[axis]);
        size_t j = i;
        while (j > 0 && total_key(v[j - 1].c[axis]) > kx) { v[j] = v[j - 1]; --j; }
        v[j] = x;
    }
}

// Stable bottom-up merge sort with caller-provided scratch (std::stable_sort allocates on every call).
void merge_sort_records(CenRec* v, CenRec* tmp, size_t n, int axis) {
    const size_t RUN = 16;
    for (size_t b = 0; b < n; b += RUN) insertion_sort_records(v + b, std::min(RUN, n - b), axis);
    CenRec* src = v;
    CenRec* dst = tmp;
    for (size_t width = RUN; width < n; width *= 2) {
        for (size_t b = 0; b < n; b += 2 * width) {
            const size_t m = std::min(b + width, n), e = std::min(b + 2 * width, n);
            size_t i = b, j = m, k = b;
            while (i < m && j < e) dst[k++] = total_key(src[j].c[axis]) < total_key(src[i].c[axis]) ? src[j++] : src[i++];
            while (i < m) dst[k++] = src[i++];
            while (j < e) dst[k++] = src[j++];
        }
        std::swap(src, dst);
    }
    if (src != v) std::memcpy(v, src, n * sizeof(CenRec));
}

struct LeafOrderJob {
    CenRec* rec;
    CenRec* tmp;  // scratch, indexed like rec: a node only uses its own range
    std::atomic<int> threads_left{0};
    LeafOrderJob(CenRec* r, CenRec* t) : rec(r), tmp(t) {}

    // bvh.rs:58-108 on the range [lo, hi): pick the axis, stable-sort the range by it. T = threads for this node.
    void sort_node(size_t lo, size_t hi, size_t T) {
        CenRec* v = rec + lo;
        const size_t len = hi - lo;
        float cmin[3], cmax[3];
        if (T > 1) {
            std::vector<float> part(T * 6);
            const size_t used = parallel_chunks(len, T, [&](size_t t, size_t b, size_t e) {
                float mn[3], mx[3];
                for (int a = 0; a < 3; ++a) mn[a] = mx[a] = v[b].c[a];
                for (size_t i = b + 1; i < e; ++i)
                    for (int a = 0; a < 3; ++a) {
                        mn[a] = fmin_(mn[a], v[i].c[a]);
                        mx[a] = fmax_(mx[a], v[i].c[a]);
                    }
                for (int a = 0; a < 3; ++a) { part[6 * t + a] = mn[a]; part[6 * t + 3 + a] = mx[a]; }
            });
            for (int a = 0; a < 3; ++a) { cmin[a] = part[a]; cmax[a] = part[3 + a]; }
            for (size_t t = 1; t < used; ++t)
                for (int a = 0; a < 3; ++a) {
                    cmin[a] = fmin_(cmin[a], part[6 * t + a]);
                    cmax[a] = fmax_(cmax[a], part[6 * t + 3 + a]);
                }
        } else {
            for (int a = 0; a < 3; ++a) cmin[a] = cmax[a] = v[0].c[a];
            for (size_t i = 1; i < len; ++i)
                for (int a = 0; a < 3; ++a) {
                    cmin[a] = fmin_(cmin[a], v[i].c[a]);
                    cmax[a] = fmax_(cmax[a], v[i].c[a]);
                }
        }
        const float sx = cmax[0] - cmin[0], sy = cmax[1] - cmin[1], sz = cmax[2] - cmin[2];
        int axis;  // bvh.rs:71-77: strictly largest spread, else Z
        if (sx > sy && sx > sz) axis = 0;
        else if (sy > sx && sy > sz) axis = 1;
        else axis = 2;
        // Vec::sort_by(total_cmp) is stable, bvh.rs:80-108. A child that keeps its parent's axis is already in
        // order (a stable sort of a sorted range is the identity): one linear check saves the sort.
        bool sorted = true;
        for (size_t i = 1; i < len && sorted; ++i) sorted = total_key(v[i - 1].c[axis]) <= total_key(v[i].c[axis]);
        if (sorted) return;
        if (len <= 24) insertion_sort_records(v, len, axis);
        else if (len < 4096) merge_sort_records(v, tmp + lo, len, axis);
        else radix_sort_records(v, tmp + lo, len, axis, T);
    }

    void run(size_t lo, size_t hi) {
        std::vector<std::pair<size_t, size_t>> stack;
        std::vector<std::thread> spawned;
        stack.emplace_back(lo, hi);
        while (!stack.empty()) {
            const std::pair<size_t, size_t> r = stack.back();
            stack.pop_back();
            const size_t len = r.second - r.first;
            if (len < 2) continue;
            sort_node(r.first, r.second, 1);
            const size_t mid = r.first + len / 2;  // bvh.rs:111-120
            // the halves are independent: hand the left one to another thread while any are free
            bool handed_over = false;
            if (len > 4096) {
                if (threads_left.fetch_sub(1) > 0)
                    handed_over = try_spawn(spawned, [this, r, mid]() {
                        run(r.first, mid);
                        threads_left.fetch_add(1);
                    });
                if (!handed_over) threads_left.fetch_add(1);
            }
            if (!handed_over) stack.emplace_back(r.first, mid);
            stack.emplace_back(mid, r.second);
        }
        for (std::thread& t : spawned) t.join();
    }
};

}  // namespace

// core/bvh.rs:48-130 restated on index ranges: a range of >= 2 items is stably sorted by the centroid
// coordinate of the axis with the strictly largest centroid spread (else Z) and cut at len/2. The left
// half precedes the right half, so after the recursion the array itself is the in-order leaf sequence.
// boxes6: n x (min xyz, max xyz).
void reference_leaf_order(const float* boxes6, size_t n, RawVector<uint32_t>& order) {
    order.resize(n);
    if (n < 2) {
        if (n) order[0] = 0;
        return;
    }
    RawVector<CenRec> rec(n), tmp(n);
    parallel_for(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            // centroid = (min + max) / 2.0 (util/aabb.rs:33-35)
            for (int a = 0; a < 3; ++a) rec[i].c[a] = (boxes6[6 * i + a] + boxes6[6 * i + 3 + a]) / 2.0f;
            rec[i].idx = (uint32_t)i;
        }
    });
    LeafOrderJob job(rec.data(), tmp.data());
    const size_t T = worker_count();
    // the few nodes at the top are sorted one after the other with every thread; what is left below them are
    // independent subtrees, one task each (a task hands halves to further threads while any are free)
    size_t PAR_MIN = (size_t)1 << 20;
    if (const char* e = std::getenv("VOIDRAY_PAR_MIN")) PAR_MIN = (size_t)std::max(2, atoi(e));  // test knob
    std::vector<std::pair<size_t, size_t>> top, tasks;
    top.emplace_back((size_t)0, n);
    while (!top.empty()) {
        const std::pair<size_t, size_t> r = top.back();
        top.pop_back();
        const size_t len = r.second - r.first;
        if (len < PAR_MIN || T == 1) {
            tasks.push_back(r);
            continue;
        }
        job.sort_node(r.first, r.second, T);
        const size_t mid = r.first + len / 2;
        top.emplace_back(r.first, mid);
        top.emplace_back(mid, r.second);
    }
    job.threads_left = (int)T - (int)tasks.size();
    if (tasks.size() == 1) {
        job.threads_left = (int)T - 1;
        job.run(tasks[0].first, tasks[0].second);
    } else {
        std::vector<std::thread> th;
        for (const std::pair<size_t, size_t>& r : tasks)
            if (!try_spawn(th, [&job, r]() { job.run(r.first, r.second); })) job.run(r.first, r.second);
        for (std::thread& t : th) t.join();
    }
    parallel_for(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) order[i] = rec[i].idx;
    });
}

namespace {

// ------------------------------------------------------------------------------------------------
// Binned-SAH BVH2 builder
// ------------------------------------------------------------------------------------------------
struct Box {
    float lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    }
    void grow(const float* p) {
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); }
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
};

// One primitive of the build: 32 bytes, two 16-byte halves (min | id, max | 0) that the passes below read with one
// SSE load each. The centroid 0.5 * (min + max) is recomputed where it is needed instead of being carried along.
struct Prim {
    float lo[3];
    uint32_t id;
    float hi[3];
    uint32_t pad;
    float cen(int a) const { return 0.5f * (lo[a] + hi[a]); }
};
static_assert(sizeof(Prim) == 32, "Prim is two 16-byte halves");

// Box in SSE registers (lanes x, y, z; lane 3 carries whatever the loads brought along and is never read).
// _mm_min_ps(p, lo) == (p < lo ? p : lo) == std::min(lo, p) and _mm_max_ps(p, hi) == std::max(hi, p) operand for
// operand, so the bounds are bit-identical to the scalar Box above, signed zeros and NaNs included.
struct SBox {
    __m128 lo, hi;
    void reset() {
        lo = _mm_set1_ps(FLT_MAX);
        hi = _mm_set1_ps(-FLT_MAX);
    }
    void grow(__m128 plo, __m128 phi) {
        lo = _mm_min_ps(plo, lo);
        hi = _mm_max_ps(phi, hi);
    }
    void grow(__m128 point) { grow(point, point); }
    void grow(const SBox& b) { grow(b.lo, b.hi); }
    void grow(const Prim& p) { grow(_mm_loadu_ps(p.lo), _mm_loadu_ps(p.hi)); }
    void grow(const Box& b) {
        grow(_mm_setr_ps(b.lo[0], b.lo[1], b.lo[2], 0.0f), _mm_setr_ps(b.hi[0], b.hi[1], b.hi[2], 0.0f));
    }
    Box box() const {
        float l[4], h[4];
        _mm_storeu_ps(l, lo);
        _mm_storeu_ps(h, hi);
        Box b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = l[a]; b.hi[a] = h[a]; }
        return b;
    }
    float half_area() const { return box().half_area(); }
};
// centroid of a primitive in lanes x, y, z (lane 3 = 0: the id bits are masked out before the arithmetic, a small
// integer read as a float is a denormal)
inline __m128 prim_centroid(const Prim& p) {
    const __m128 mask = _mm_castsi128_ps(_mm_setr_epi32(-1, -1, -1, 0));
    return _mm_mul_ps(_mm_set1_ps(0.5f), _mm_add_ps(_mm_and_ps(_mm_loadu_ps(p.lo), mask), _mm_loadu_ps(p.hi)));
}

struct Builder {
    RawVector<Prim>& prims;
    RawVector<Quad>& nodes;
    std::atomic<uint32_t> next_node{0};
    std::atomic<uint32_t> max_depth{0};
    std::atomic<int> threads_left{0};
    static const int BINS = 32;
    static const uint32_t MAX_DEPTH = 32;  // == the kernel's shared-memory stack depth
    uint32_t leaf_max = LEAF_MAX_TRIS;
    float node_cost = 1.0f;

    RawVector<uint32_t> subtree_nodes;  // per node: inner nodes in its subtree, itself included (for the renumbering)
    Builder(RawVector<Prim>& p, RawVector<Quad>& n) : prims(p), nodes(n) { subtree_nodes.resize(n.size() / NODE_QUADS); }

    // Copies the subtree rooted at old node `from` into `dst` in depth-first pre-order starting at index `to`
    // (parent, left subtree, right subtree): the new index of a right child is known from the left subtree's size,
    // so large subtrees are renumbered on other threads.
    void renumber(uint32_t from, uint32_t to, RawVector<Quad>& dst) {
        struct Item {
            uint32_t from, to;
        };
        std::vector<Item> todo;
        std::vector<std::thread> spawned;
        todo.push_back(Item{from, to});
        while (!todo.empty()) {
            const Item it = todo.back();
            todo.pop_back();
            const Quad* src = &nodes[(size_t)it.from * NODE_QUADS];
            Quad* out = &dst[(size_t)it.to * NODE_QUADS];
            int32_t c0, c1;
            std::memcpy(&c0, &src[1].z, 4);
            std::memcpy(&c1, &src[1].w, 4);
            const uint32_t left_to = it.to + 1, right_to = it.to + 1 + (c0 >= 0 ? subtree_nodes[c0] : 0u);
            out[0] = src[0];
            out[1] = Quad{src[1].x, src[1].y, c0 >= 0 ? bits_f(left_to) : src[1].z, c1 >= 0 ? bits_f(right_to) : src[1].w};
            if (c1 >= 0) {
                bool handed_over = false;
                if (subtree_nodes[c1] > 65536) {
                    if (threads_left.fetch_sub(1) > 0)
                        handed_over = try_spawn(spawned, [this, c1, right_to, &dst]() {
                            renumber((uint32_t)c1, right_to, dst);
                            threads_left.fetch_add(1);
                        });
                    if (!handed_over) threads_left.fetch_add(1);
                }
                if (!handed_over) todo.push_back(Item{(uint32_t)c1, right_to});
            }
            if (c0 >= 0) todo.push_back(Item{(uint32_t)c0, left_to});
        }
        for (std::thread& t : spawned) t.join();
    }

    struct Bins {
        SBox box[3][BINS];
        uint32_t count[3][BINS];
        void reset() {
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < BINS; ++b) { box[a][b].reset(); count[a][b] = 0; }
        }
        void merge(const Bins& o) {
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < BINS; ++b) { box[a][b].grow(o.box[a][b]); count[a][b] += o.count[a][b]; }
        }
    };

    // Pieces a large node is cut into for its bounds / binning / partition passes. A function of the node size
    // only, not of the machine, so the tree is the same wherever it is built.
    static size_t par_chunks(size_t n) { return std::min<size_t>(std::max<size_t>(n >> 18, 1), 32); }

    // In-place partition of v[0, n) by pred(piece, element) in `chunks` pieces: each piece is partitioned on its own thread, then the
    // right-side elements left of the global split are swapped with the left-side elements right of it.
    template <typename Pred>
    static size_t parallel_partition(Prim* v, size_t n, size_t chunks, Pred pred) {
        std::vector<size_t> cb(chunks, 0), cm(chunks, 0), ce(chunks, 0);
        const size_t used = parallel_chunks(n, chunks, [&](size_t t, size_t b, size_t e) {
            cb[t] = b;
            ce[t] = e;
            cm[t] = (size_t)(std::partition(v + b, v + e, [&pred, t](const Prim& p) { return pred(t, p); }) - v);
        });
        size_t mid = 0;
        for (size_t t = 0; t < used; ++t) mid += cm[t] - cb[t];
        struct Range {
            size_t at, len;
        };
        std::vector<Range> wrong_left, wrong_right;  // right-side elements in [0, mid), left-side ones in [mid, n)
        for (size_t t = 0; t < used; ++t) {
            const size_t rb = cm[t], re = std::min(ce[t], mid);
            if (rb < re) wrong_left.push_back(Range{rb, re - rb});
            const size_t lb = std::max(cb[t], mid), le = cm[t];
            if (lb < le) wrong_right.push_back(Range{lb, le - lb});
        }
        struct Swap {
            size_t a, b, len;
        };
        std::vector<Swap> swaps;
        size_t i = 0, j = 0;
        const size_t PIECE = 1 << 16;
        while (i < wrong_left.size() && j < wrong_right.size()) {
            const size_t len = std::min(std::min(wrong_left[i].len, wrong_right[j].len), PIECE);
            swaps.push_back(Swap{wrong_left[i].at, wrong_right[j].at, len});
            wrong_left[i].at += len;
            wrong_left[i].len -= len;
            wrong_right[j].at += len;
            wrong_right[j].len -= len;
            if (!wrong_left[i].len) ++i;
            if (!wrong_right[j].len) ++j;
        }
        parallel_chunks(swaps.size(), chunks, [&](size_t, size_t b, size_t e) {
            for (size_t s = b; s < e; ++s) std::swap_ranges(v + swaps[s].a, v + swaps[s].a + swaps[s].len, v + swaps[s].b);
        });
        return mid;
    }

    static int32_t leaf_code(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | count); }

    void note_depth(uint32_t d) {
        uint32_t cur = max_depth.load();
        while (d > cur && !max_depth.compare_exchange_weak(cur, d)) {}
    }

    // Builds the subtree over prims[lo, hi); returns its child code and box. known: the range's bounds and
    // centroid bounds when the parent's partition pass has already gathered them.
    int32_t build(uint32_t lo, uint32_t hi, Box& out_box, uint32_t depth, const Box* known = nullptr) {
        const uint32_t n = hi - lo;
        Box bounds, cbounds;
        bounds.reset();
        cbounds.reset();
        if (known) {
            bounds = known[0];
            cbounds = known[1];
        } else if (par_chunks(n) > 1) {
            std::vector<Box> part(2 * par_chunks(n));
            const size_t used = parallel_chunks(n, par_chunks(n), [&](size_t t, size_t b, size_t e) {
                SBox bb, cb;
                bb.reset();
                cb.reset();
                for (size_t i = lo + b; i < lo + e; ++i) {
                    bb.grow(prims[i]);
                    cb.grow(prim_centroid(prims[i]));
                }
                part[2 * t] = bb.box();
                part[2 * t + 1] = cb.box();
            });
            for (size_t t = 0; t < used; ++t) {
                bounds.grow(part[2 * t]);
                cbounds.grow(part[2 * t + 1]);
            }
        } else {
            SBox bb, cb;
            bb.reset();
            cb.reset();
            for (uint32_t i = lo; i < hi; ++i) {
                bb.grow(prims[i]);
                cb.grow(prim_centroid(prims[i]));
            }
            bounds = bb.box();
            cbounds = cb.box();
        }
        out_box = bounds;
        if (n == 1) {
            note_depth(depth);
            return leaf_code(lo, n);
        }

        float best_cost = FLT_MAX;
        int best_axis = -1, best_bin = -1;
        const float parent_area = std::max(bounds.half_area(), 1e-30f);

        // small ranges (most of the nodes): exact sweep SAH over the sorted centroids instead of 32 bins
        static const uint32_t SWEEP_MAX = 16;
        uint32_t sweep_left = 0;  // > 0: the range has been reordered and the split is after this many prims
        if (n <= SWEEP_MAX) {
            uint8_t order[SWEEP_MAX], best_order[SWEEP_MAX];
            float right_area[SWEEP_MAX];
            for (int axis = 0; axis < 3; ++axis) {
                if (!(cbounds.hi[axis] > cbounds.lo[axis])) continue;
                float cen[SWEEP_MAX];
                for (uint32_t i = 0; i < n; ++i) {
                    order[i] = (uint8_t)i;
                    cen[i] = prims[lo + i].cen(axis);
                }
                for (uint32_t i = 1; i < n; ++i) {  // insertion sort by centroid
                    const uint8_t v = order[i];
                    const float key = cen[v];
                    uint32_t j = i;
                    while (j > 0 && cen[order[j - 1]] > key) { order[j] = order[j - 1]; --j; }
                    order[j] = v;
                }
                SBox acc;
                acc.reset();
                for (uint32_t i = n - 1; i > 0; --i) {
                    acc.grow(prims[lo + order[i]]);
                    right_area[i] = acc.half_area();
                }
                acc.reset();
                for (uint32_t i = 1; i < n; ++i) {  // split: the first i prims go left
                    acc.grow(prims[lo + order[i - 1]]);
                    const float cost = (acc.half_area() * (float)i + right_area[i] * (float)(n - i)) / parent_area;
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        sweep_left = i;
                        std::memcpy(best_order, order, n);
                    }
                }
            }
            if (best_axis >= 0 && !(n <= leaf_max && (float)n <= node_cost + best_cost)) {
                Prim tmp[SWEEP_MAX];
                for (uint32_t i = 0; i < n; ++i) tmp[i] = prims[lo + best_order[i]];
                for (uint32_t i = 0; i < n; ++i) prims[lo + i] = tmp[i];
            }
        }

        // binned SAH over the three axes, one pass over the primitives (split over threads for a large node:
        // the bins are merged with min / max / +, so the result does not depend on the chunking)
        const size_t chunks = par_chunks(n);
        if (n > SWEEP_MAX) {
            float cmin[3], k[3];
            bool valid[3];
            for (int axis = 0; axis < 3; ++axis) {
                cmin[axis] = cbounds.lo[axis];
                valid[axis] = cbounds.hi[axis] > cbounds.lo[axis];
                k[axis] = valid[axis] ? (float)BINS * (1.0f - 1e-6f) / (cbounds.hi[axis] - cbounds.lo[axis]) : 0.0f;
            }
            // all three bin indices of a primitive at once: (int)((centroid - cmin) * k) per lane (cvttps2dq truncates
            // like the scalar conversion and gives INT_MIN out of range, which the clamp sends to bin 0 either way);
            // a flat axis (k = 0) lands in bin 0 and is skipped by the sweep below
            const __m128 cmin4 = _mm_setr_ps(cmin[0], cmin[1], cmin[2], 0.0f), k4 = _mm_setr_ps(k[0], k[1], k[2], 0.0f);
            auto bin_range = [&](uint32_t b0, uint32_t e0, Bins& bins) {
                bins.reset();
                for (uint32_t i = b0; i < e0; ++i) {
                    const Prim& p = prims[i];
                    const __m128 plo = _mm_loadu_ps(p.lo), phi = _mm_loadu_ps(p.hi);
                    alignas(16) int bi[4];
                    _mm_store_si128((__m128i*)bi, _mm_cvttps_epi32(_mm_mul_ps(_mm_sub_ps(prim_centroid(p), cmin4), k4)));
                    for (int axis = 0; axis < 3; ++axis) {
                        const int b = std::min(std::max(bi[axis], 0), BINS - 1);
                        bins.box[axis][b].grow(plo, phi);
                        bins.count[axis][b]++;
                    }
                }
            };
            Bins bins;
            if (chunks > 1) {
                std::vector<Bins> part(chunks);
                const size_t used = parallel_chunks(n, chunks, [&](size_t t, size_t b, size_t e) {
                    bin_range(lo + (uint32_t)b, lo + (uint32_t)e, part[t]);
                });
                bins = part[0];
                for (size_t t = 1; t < used; ++t) bins.merge(part[t]);
            } else {
                bin_range(lo, hi, bins);
            }
            for (int axis = 0; axis < 3; ++axis) {
                if (!valid[axis]) continue;
                float right_area[BINS];
                SBox acc;
                acc.reset();
                for (int b = BINS - 1; b > 0; --b) {
                    acc.grow(bins.box[axis][b]);
                    right_area[b] = acc.half_area();
                }
                acc.reset();
                uint32_t left_n = 0;
                for (int b = 0; b < BINS - 1; ++b) {
                    acc.grow(bins.box[axis][b]);
                    left_n += bins.count[axis][b];
                    if (left_n == 0 || left_n == n) continue;
                    const float cost = (acc.half_area() * (float)left_n + right_area[b + 1] * (float)(n - left_n)) / parent_area;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
        }

        // SAH in units of one triangle test: a node visit (two slab tests + stack work) costs node_cost
        if (n <= leaf_max && (best_axis < 0 || (float)n <= node_cost + best_cost)) {
            note_depth(depth);
            return leaf_code(lo, n);
        }

        // depth cap: the traversal stack holds MAX_DEPTH entries. Once the remaining levels only just suffice
        // for a balanced tree over n primitives, split at the median (by position) instead of by SAH.
        uint32_t levels_needed = 0;
        while ((1u << levels_needed) < n) ++levels_needed;
        const bool force_median = depth + levels_needed + 1 >= MAX_DEPTH;
        uint32_t mid;
        Box kids[4];  // left bounds, left centroid bounds, right bounds, right centroid bounds
        bool have_kids = false;
        if (n <= SWEEP_MAX && best_axis >= 0) {
            mid = force_median ? lo + n / 2 : lo + sweep_left;  // the range is already sorted along best_axis
        } else if (force_median && best_axis >= 0) {
            Prim* first = prims.data() + lo;
            Prim* nth = first + n / 2;
            const int ax = best_axis;
            std::nth_element(first, nth, prims.data() + hi, [ax](const Prim& a, const Prim& b) { return a.cen(ax) < b.cen(ax); });
            mid = lo + n / 2;
        } else if (best_axis < 0) {
            mid = lo + n / 2;  // identical centroids: split by position
        } else {
            const float cmin = cbounds.lo[best_axis], cmax = cbounds.hi[best_axis];
            const float k = (float)BINS * (1.0f - 1e-6f) / (cmax - cmin);
            auto goes_left = [&](const Prim& p) {
                int b = (int)((p.cen(best_axis) - cmin) * k);
                b = std::min(std::max(b, 0), BINS - 1);
                return b <= best_bin;
            };
            // the partition pass also gathers both sides' bounds (the predicate runs exactly once per element)
            for (int i = 0; i < 4; ++i) kids[i].reset();
            if (chunks > 1) {
                std::vector<SBox> part(4 * chunks);
                for (SBox& b : part) b.reset();
                mid = lo + (uint32_t)parallel_partition(prims.data() + lo, n, chunks, [&](size_t t, const Prim& p) {
                    const bool left = goes_left(p);
                    SBox* kb = &part[4 * t + (left ? 0 : 2)];
                    kb[0].grow(p);
                    kb[1].grow(prim_centroid(p));
                    return left;
                });
                for (size_t t = 0; t < chunks; ++t)
                    for (int i = 0; i < 4; ++i) kids[i].grow(part[4 * t + i].box());
            } else {
                SBox sk[4];
                for (int i = 0; i < 4; ++i) sk[i].reset();
                Prim* m = std::partition(prims.data() + lo, prims.data() + hi, [&](const Prim& p) {
                    const bool left = goes_left(p);
                    SBox* kb = sk + (left ? 0 : 2);
                    kb[0].grow(p);
                    kb[1].grow(prim_centroid(p));
                    return left;
                });
                for (int i = 0; i < 4; ++i) kids[i] = sk[i].box();
                mid = (uint32_t)(m - prims.data());
            }
            have_kids = true;
            if (mid == lo || mid == hi) {
                mid = lo + n / 2;
                have_kids = false;
            }
        }

        const uint32_t node = next_node.fetch_add(1);
        Box lbox, rbox;
        int32_t lcode, rcode;
        bool spawned = false;
        if (n > 2048) {
            if (threads_left.fetch_sub(1) > 0) {
                std::vector<std::thread> left_thread;
                spawned = try_spawn(left_thread, [&]() { lcode = build(lo, mid, lbox, depth + 1, have_kids ? kids : nullptr); });
                if (spawned) {
                    rcode = build(mid, hi, rbox, depth + 1, have_kids ? kids + 2 : nullptr);
                    left_thread[0].join();
                }
            }
            threads_left.fetch_add(1);
        }
        if (!spawned) {
            lcode = build(lo, mid, lbox, depth + 1, have_kids ? kids : nullptr);
            rcode = build(mid, hi, rbox, depth + 1, have_kids ? kids + 2 : nullptr);
        }
        write_node(node, lcode, lbox, rcode, rbox);
        subtree_nodes[node] = 1u + (lcode >= 0 ? subtree_nodes[lcode] : 0u) + (rcode >= 0 ? subtree_nodes[rcode] : 0u);
        return (int32_t)node;
    }

    // Quantisation grid (layout.h): 15-bit cells over the scene bounds, set before the build.
    double grid_min[3] = {0, 0, 0}, grid_inv_cell[3] = {1, 1, 1};

    // lo rounded down, hi rounded up, one more cell outwards for the decoder's rounding; an empty box
    // (lo > hi) becomes the inverted pair (32767, 0), which no ray enters.
    uint32_t quantize_pair(float lo, float hi, int axis) const {
        if (!(lo <= hi)) return (0x8000u | 32767u) | ((0x8000u | 0u) << 16);
        // clamped as doubles (infinite or NaN coordinates must not reach the integer conversion); a NaN cell
        // index opens the box to the whole grid, the conservative side
        double fl = std::floor(((double)lo - grid_min[axis]) * grid_inv_cell[axis]) - 1.0;
        double fh = std::ceil(((double)hi - grid_min[axis]) * grid_inv_cell[axis]) + 1.0;
        if (!(fl >= 0.0)) fl = 0.0;
        if (fl > 32767.0) fl = 32767.0;
        if (!(fh <= 32767.0)) fh = 32767.0;
        if (fh < 0.0) fh = 0.0;
        return (0x8000u | (uint32_t)fl) | ((0x8000u | (uint32_t)fh) << 16);
    }

    void write_node(uint32_t node, int32_t c0, const Box& b0, int32_t c1, const Box& b1) {
        Quad* q = nodes.data() + (size_t)node * NODE_QUADS;
        q[0] = Quad{bits_f(quantize_pair(b0.lo[0], b0.hi[0], 0)), bits_f(quantize_pair(b0.lo[1], b0.hi[1], 1)),
                    bits_f(quantize_pair(b0.lo[2], b0.hi[2], 2)), bits_f(quantize_pair(b1.lo[0], b1.hi[0], 0))};
        q[1] = Quad{bits_f(quantize_pair(b1.lo[1], b1.hi[1], 1)), bits_f(quantize_pair(b1.lo[2], b1.hi[2], 2)),
                    bits_f((uint32_t)c0), bits_f((uint32_t)c1)};
    }
};


}  // namespace

// A texture whose every value is exactly v / 255 (an 8-bit source through to_rgb32f) also keeps its
// RGBA8 form; the kernel's conversion reproduces the same floats, so the lookups do not change by a bit.
void pack_texture_rgba8(HostTexture& t) {
    const size_t n = (size_t)t.w * t.h;
    std::vector<uint8_t> out(4 * n);
    for (size_t i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) {
            const float f = t.rgb[3 * i + c];
            const float scaled = f * 255.0f;
            if (!(scaled >= 0.0f && scaled <= 255.0f)) return;
            const int v = (int)(scaled + 0.5f);
            const float back = (float)v / 255.0f;
            if (std::memcmp(&back, &f, 4) != 0) return;  // bit for bit (-0.0 is not an 8-bit value)
            out[4 * i + c] = (uint8_t)v;
        }
        out[4 * i + 3] = 0;
    }
    t.rgba8.swap(out);
}

namespace {
// VOIDRAY_TIMING=1 prints the host phases of a commit to stderr
struct PhaseTimer {
    bool on = std::getenv("VOIDRAY_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[voidray] flatten: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
}  // namespace

bool flatten_scene(const HostScene& in, FlatScene& out, std::string& err) {
    out = FlatScene();
    PhaseTimer timer;
    const size_t n_surfaces = in.surfaces.size();
    // scene.rs:183-184 resolves a hit's material through objects[surface index]
    if (in.objects.size() < n_surfaces) {
        err = "every surface needs an object at its own index: the reference looks the material of surface s up as "
              "objects[s] (core/scene.rs:183-184) and would panic on the first hit of surface " +
              std::to_string(in.objects.size());
        return false;
    }
    for (size_t s = 0; s < in.objects.size(); ++s) {
        if (in.objects[s].material >= in.materials.size()) { err = "object refers to an unknown material"; return false; }
        if (in.objects[s].surface >= n_surfaces) { err = "object refers to an unknown surface"; return false; }
    }
    for (const MaterialRec& m : in.materials) {
        if (m.albedo_tex >= (int32_t)in.textures.size() || m.normal_tex >= (int32_t)in.textures.size()) {
            err = "material refers to an unknown texture";
            return false;
        }
    }

    // ---- camera (core/camera.rs:38-54) ----
    {
        const HostCamera& c = in.camera;
        CameraRec& r = out.camera;
        const V3 dir = ld3(c.direction), up = ld3(c.up), eye = ld3(c.eye);
        r.d = 1.0f / std::tan(c.fov / 2.0f);
        const V3 right = normalize(cross(dir, up));
        std::memcpy(r.origin, c.eye, 12);
        std::memcpy(r.direction, c.direction, 12);
        std::memcpy(r.up, c.up, 12);
        r.right[0] = right.x; r.right[1] = right.y; r.right[2] = right.z;
        r.has_dof = c.has_dof;
        r.aperture = c.aperture;
        r.focal_length = c.has_dof ? dot(sub(ld3(c.focal_point), eye), dir) : 0.0f;
    }

    // ---- per-surface bounds (scene.rs:72-79) ----
    std::vector<float> surface_boxes(6 * n_surfaces);
    out.mesh_tie_rank.resize(n_surfaces);
    uint64_t total_tris = 0;
    for (size_t s = 0; s < n_surfaces; ++s) {
        const HostSurface& sf = in.surfaces[s];
        float* b = &surface_boxes[6 * s];
        if (sf.kind == 0) {
            const HostMesh& m = in.meshes[sf.mesh];
            // mesh.rs:93-102: AABB::default() grown by every vertex
            b[0] = b[1] = b[2] = INFINITY;
            b[3] = b[4] = b[5] = -INFINITY;
            for (uint32_t v = 0; v < m.n_vertices; ++v)
                for (int a = 0; a < 3; ++a) {
                    b[a] = fmin_(b[a], m.pos[3 * v + a]);
                    b[3 + a] = fmax_(b[3 + a], m.pos[3 * v + a]);
                }
            total_tris += m.idx.size() / 3;
        } else if (sf.kind == 1) {  // surfaces.rs:36-43
            for (int a = 0; a < 3; ++a) {
                b[a] = sf.center[a] - sf.radius_or_height;
                b[3 + a] = sf.center[a] + sf.radius_or_height;
            }
        } else {  // surfaces.rs:107-114
            b[0] = -INFINITY; b[1] = sf.radius_or_height - 0.0001f; b[2] = -INFINITY;
            b[3] = INFINITY; b[4] = sf.radius_or_height + 0.0001f; b[5] = INFINITY;
        }
    }
    timer.lap("surface bounds");
    if (total_tris >= (1u << 28)) { err = "too many triangles (limit 2^28)"; return false; }

    // ---- per-mesh tie ranks: the reference tree's in-order leaf sequence (mesh.rs:112-141) ----
    // Only the packed triangle records at the end read them, so they are computed on a second thread while this
    // one builds the BVH (the two share nothing; both fan out further on their own).
    auto tie_ranks = [&in, &out, n_surfaces]() {
        PhaseTimer rank_timer;
        for (size_t s = 0; s < n_surfaces; ++s) {
            const HostSurface& sf = in.surfaces[s];
            if (sf.kind != 0) continue;
            const HostMesh& m = in.meshes[sf.mesh];
            const size_t nt = m.idx.size() / 3;
            RawVector<uint32_t>& rank = out.mesh_tie_rank[s];
            rank.resize(nt);
            if (m.idx.size() > 4 * 3) {  // SMALL_MESH, mesh.rs:43,112
                // per-triangle boxes, mesh.rs:193-205: vertex bounds, epsilon_expand(0.001)
                RawVector<float> boxes(6 * nt);
                parallel_for(nt, [&](size_t t_begin, size_t t_end) {
                for (size_t t = t_begin; t < t_end; ++t) {
                    float* tb = &boxes[6 * t];
                    for (int a = 0; a < 3; ++a) {
                        const float p0 = m.pos[3 * m.idx[3 * t] + a], p1 = m.pos[3 * m.idx[3 * t + 1] + a],
                                    p2 = m.pos[3 * m.idx[3 * t + 2] + a];
                        tb[a] = fmin_(p0, fmin_(p1, p2));
                        tb[3 + a] = fmax_(p0, fmax_(p1, p2));
                    }
                    for (int a = 0; a < 3; ++a) {  // util/aabb.rs:62-83
                        const float dim = tb[3 + a] - tb[a];
                        const float c = (tb[a] + tb[3 + a]) / 2.0f;
                        if (dim < 0.001f) { tb[a] = c - 0.001f; tb[3 + a] = c + 0.001f; }
                    }
                }
                });
                rank_timer.lap("(2nd thread) triangle boxes");
                RawVector<uint32_t> order;
                reference_leaf_order(boxes.data(), nt, order);
                parallel_for(nt, [&](size_t i_begin, size_t i_end) {
                    for (size_t i = i_begin; i < i_end; ++i) rank[order[i]] = (uint32_t)i;
                });
                rank_timer.lap("(2nd thread) ref. leaf order");
            } else {
                // linear loop, first index wins a tie (mesh.rs:129-135)
                for (size_t t = 0; t < nt; ++t) rank[t] = (uint32_t)(nt - 1 - t);
            }
        }
    };
    std::vector<std::thread> rank_thread;
    std::exception_ptr rank_error;  // an exception of the second thread (bad_alloc) is rethrown on this one
    auto tie_ranks_guarded = [&tie_ranks, &rank_error]() {
        try {
            tie_ranks();
        } catch (...) {
            rank_error = std::current_exception();
        }
    };
    if (total_tris < 1024 || std::thread::hardware_concurrency() < 2 || !try_spawn(rank_thread, tie_ranks_guarded)) tie_ranks();
    struct Joiner {  // joined before the records are packed, and on every early return
        std::vector<std::thread>& t;
        void join() { for (std::thread& x : t) if (x.joinable()) x.join(); }
        ~Joiner() { join(); }
    } rank_join{rank_thread};

    // ---- scene-level in-order sequence -> global rank base of every surface ----
    RawVector<uint32_t> surface_order;
    reference_leaf_order(surface_boxes.data(), n_surfaces, surface_order);
    out.surface_rank_base.assign(n_surfaces, 0);
    {
        uint32_t base = 0;
        for (size_t i = 0; i < n_surfaces; ++i) {
            const uint32_t s = surface_order[i];
            out.surface_rank_base[s] = base;
            const HostSurface& sf = in.surfaces[s];
            base += sf.kind == 0 ? (uint32_t)(in.meshes[sf.mesh].idx.size() / 3) : 1u;
        }
    }

    // ---- the reference's scene-level tree, kept for its culling semantics (layout.h: SceneTreeNode) ----
    // from_list cuts a sorted range at len / 2 (bvh.rs:111-120), so the in-order sequence determines the tree: a range
    // of >= 2 surfaces is a Split whose box is the union of their bounds (bvh.rs:57-61), a single surface an Object.
    out.scene_tree.clear();
    out.surface_node.assign(n_surfaces, 0);
    if (n_surfaces >= 2) {
        out.scene_tree.reserve(2 * n_surfaces - 1);
        struct Range {
            size_t b, e;
            int32_t parent;
        };
        std::vector<Range> todo;
        todo.push_back(Range{0, n_surfaces, -1});
        while (!todo.empty()) {  // pre-order: left subtree before right
            const Range r = todo.back();
            todo.pop_back();
            SceneTreeNode node;
            node.parent = r.parent;
            const int32_t self = (int32_t)out.scene_tree.size();
            if (r.e - r.b == 1) {
                const uint32_t s = surface_order[r.b];
                for (int a = 0; a < 3; ++a) { node.lo[a] = 0.0f; node.hi[a] = 0.0f; }
                node.a = ~(int32_t)s;
                out.surface_node[s] = (uint32_t)self;
            } else {
                for (int a = 0; a < 3; ++a) { node.lo[a] = INFINITY; node.hi[a] = -INFINITY; }
                for (size_t i = r.b; i < r.e; ++i) {
                    const float* b = &surface_boxes[6 * surface_order[i]];
                    for (int a = 0; a < 3; ++a) {  // AABB::surround, util/aabb.rs:46-59
                        node.lo[a] = fmin_(node.lo[a], b[a]);
                        node.hi[a] = fmax_(node.hi[a], b[3 + a]);
                    }
                }
                node.a = self + (int32_t)(2 * (r.e - r.b) - 1);  // a subtree over k surfaces has 2k - 1 nodes
                const size_t mid = r.b + (r.e - r.b) / 2;
                todo.push_back(Range{mid, r.e, self});
                todo.push_back(Range{r.b, mid, self});
            }
            out.scene_tree.push_back(node);
        }
    }

    // ---- analytic surfaces ----
    for (size_t s = 0; s < n_surfaces; ++s) {
        const HostSurface& sf = in.surfaces[s];
        if (sf.kind == 0) continue;
        AnalyticRec a;
        a.kind = sf.kind == 1 ? 0 : 1;
        a.cx = sf.center[0]; a.cy = sf.center[1]; a.cz = sf.center[2];
        a.radius = sf.radius_or_height;
        a.rank = out.surface_rank_base[s];
        a.material = in.objects[s].material;
        a.surface = (uint32_t)s;
        out.analytics.push_back(a);
    }

    // ---- triangles ----
    const uint32_t n_tris = (uint32_t)total_tris;
    out.n_tris = n_tris;
    RawVector<Prim> prims(n_tris);
    struct Src {
        uint32_t surface, prim;
    };
    RawVector<Src> src(n_tris);
    {
        uint32_t g = 0;
        for (size_t s = 0; s < n_surfaces; ++s) {
            const HostSurface& sf = in.surfaces[s];
            if (sf.kind != 0) continue;
            const HostMesh& m = in.meshes[sf.mesh];
            const size_t nt = m.idx.size() / 3;
            const uint32_t g0 = g;
            parallel_for(nt, [&](size_t t_begin, size_t t_end) {
                for (size_t t = t_begin; t < t_end; ++t) {
                    const uint32_t gi = g0 + (uint32_t)t;
                    Prim& p = prims[gi];
                    Box b;
                    b.reset();
                    for (int k = 0; k < 3; ++k) b.grow(&m.pos[3 * m.idx[3 * t + k]]);
                    for (int a = 0; a < 3; ++a) { p.lo[a] = b.lo[a]; p.hi[a] = b.hi[a]; }
                    p.id = gi;
                    p.pad = 0;
                    src[gi] = Src{(uint32_t)s, (uint32_t)t};
                }
            });
            g += (uint32_t)nt;
        }
    }

    timer.lap("primitive records");
    // ---- BVH ----
    out.nodes.resize((size_t)std::max<uint32_t>(n_tris, 2) * NODE_QUADS);
    Builder builder(prims, out.nodes);
    {
        // quantisation grid = bounds of all triangles; a flat axis gets a token extent so the cell size is not 0
        Box scene_box;
        scene_box.reset();
        for (uint32_t i = 0; i < n_tris; ++i) {
            scene_box.grow(prims[i].lo);
            scene_box.grow(prims[i].hi);
        }
        float max_extent = 0.0f;
        for (int a = 0; a < 3; ++a) max_extent = std::max(max_extent, n_tris ? scene_box.hi[a] - scene_box.lo[a] : 0.0f);
        if (!(max_extent > 0.0f)) max_extent = 1.0f;
        for (int a = 0; a < 3; ++a) {
            const float lo = n_tris ? scene_box.lo[a] : 0.0f;
            float extent = n_tris ? scene_box.hi[a] - scene_box.lo[a] : 0.0f;
            extent = std::max(extent, 1e-6f * max_extent) * 1.0001f;
            out.grid_min[a] = lo;
            out.grid_extent[a] = extent;
            builder.grid_min[a] = (double)lo;
            builder.grid_inv_cell[a] = 32768.0 / (double)extent;
        }
    }
    // tuning knobs for experiments (defaults are the shipped configuration)
    if (const char* e = std::getenv("VOIDRAY_LEAF_MAX")) builder.leaf_max = (uint32_t)std::min(7, std::max(1, atoi(e)));
    if (const char* e = std::getenv("VOIDRAY_NODE_COST")) builder.node_cost = (float)atof(e);
    Box empty;
    empty.reset();
    if (n_tris == 0) {
        builder.next_node = 1;
        builder.write_node(0, Builder::leaf_code(0, 0), empty, Builder::leaf_code(0, 0), empty);
    } else {
        builder.threads_left = (int)std::max(1u, std::thread::hardware_concurrency()) - 1;
        builder.next_node = 1;  // reserve the root
        // build the root by hand so that it is node 0 even when the whole scene fits one leaf
        Box root_box;
        // Temporarily build into a subtree; if it returns an inner node we move it to slot 0.
        const int32_t code = builder.build(0, n_tris, root_box, 1);
        if (code < 0) {
            builder.write_node(0, code, root_box, Builder::leaf_code(0, 0), empty);
        } else {
            // Subtrees built on other threads take their node numbers in whatever order the threads run. Renumber in
            // depth-first pre-order (parent, left subtree, right subtree — what a single thread produces): the layout
            // the kernel sees is then the same on every run and machine, and a parent sits next to its left child.
            const size_t n_nodes = builder.next_node.load();
            RawVector<Quad> ordered(n_nodes * NODE_QUADS);
            const uint32_t next = 1u + builder.subtree_nodes[code];
            builder.renumber((uint32_t)code, 1u, ordered);
            // node 0 is the root: a copy of node 1 (children indices are absolute, so this is a plain copy)
            for (int q = 0; q < NODE_QUADS; ++q) ordered[q] = ordered[NODE_QUADS + q];
            out.nodes.swap(ordered);
            builder.next_node = next;
        }
    }
    out.nodes.resize((size_t)builder.next_node.load() * NODE_QUADS);
    out.bvh_depth = builder.max_depth.load();
    timer.lap("SAH build");

    rank_join.join();
    if (rank_error) std::rethrow_exception(rank_error);
    timer.lap("wait for the tie ranks");
    // ---- packed records in BVH leaf order ----
    out.tri_isect.resize((size_t)n_tris * TRI_ISECT_QUADS);
    out.tri_shade.resize((size_t)n_tris * TRI_SHADE_QUADS);
    out.tri_surface.resize(n_tris);
    out.tri_prim.resize(n_tris);
    parallel_for(n_tris, [&](size_t i_begin, size_t i_end) {
    for (size_t i = i_begin; i < i_end; ++i) {
        const Src sp = src[prims[i].id];
        const HostMesh& m = in.meshes[in.surfaces[sp.surface].mesh];
        const uint32_t i0 = m.idx[3 * sp.prim], i1 = m.idx[3 * sp.prim + 1], i2 = m.idx[3 * sp.prim + 2];
        const V3 p0 = ld3(&m.pos[3 * i0]), p1 = ld3(&m.pos[3 * i1]), p2 = ld3(&m.pos[3 * i2]);
        const V3 e1 = sub(p1, p0), e2 = sub(p2, p0);  // mesh.rs:150-151
        const uint32_t rank = out.surface_rank_base[sp.surface] + out.mesh_tie_rank[sp.surface][sp.prim];
        Quad* qi = &out.tri_isect[(size_t)i * TRI_ISECT_QUADS];
        qi[0] = Quad{p0.x, p0.y, p0.z, bits_f(rank)};
        qi[1] = Quad{e1.x, e1.y, e1.z, 0.0f};
        qi[2] = Quad{e2.x, e2.y, e2.z, 0.0f};
        if (TRI_ISECT_QUADS > 3) qi[3] = Quad{0.0f, 0.0f, 0.0f, bits_f(sp.surface)};
        // geometric normal, mesh.rs:80-84
        const V3 ng = normalize(cross(sub(p2, p1), sub(p0, p1)));
        const V3 n0 = ld3(&m.nrm[3 * i0]), n1 = ld3(&m.nrm[3 * i1]), n2 = ld3(&m.nrm[3 * i2]);
        const float* t0 = &m.uv[2 * i0];
        const float* t1 = &m.uv[2 * i1];
        const float* t2 = &m.uv[2 * i2];
        Quad* qs = &out.tri_shade[(size_t)i * TRI_SHADE_QUADS];
        qs[0] = Quad{n0.x, n0.y, n0.z, t0[0]};
        qs[1] = Quad{n1.x, n1.y, n1.z, t0[1]};
        qs[2] = Quad{n2.x, n2.y, n2.z, t1[0]};
        qs[3] = Quad{ng.x, ng.y, ng.z, t1[1]};
        qs[4] = Quad{t2[0], t2[1], bits_f(in.objects[sp.surface].material), 0.0f};
        out.tri_surface[i] = sp.surface;
        out.tri_prim[i] = sp.prim;
    }
    });
    timer.lap("packed triangle records");
    return true;
}

// A texel weighs (luminance + 5 % of the mean luminance) x sin(polar angle of its row centre), with the polar
// angle of row j taken from the lookup's own mapping y = (H - 1) - acos(-d.y) / pi * H
// (voidray_common/src/environments.rs:80-86); the last row maps to no direction and gets weight 0.
void build_env_tables(const HostTexture& env, std::vector<float>& marginal, std::vector<float>& cond) {
    const size_t W = env.w, H = env.h;
    std::vector<double> weight(W * H), row_sum(H, 0.0);
    double mean = 0.0;
    for (size_t k = 0; k < W * H; ++k) {
        weight[k] = 0.2126 * (double)env.rgb[3 * k] + 0.7152 * (double)env.rgb[3 * k + 1] + 0.0722 * (double)env.rgb[3 * k + 2];
        mean += weight[k];
    }
    mean /= (double)(W * H);
    double total = 0.0;
    for (size_t j = 0; j < H; ++j) {
        const double sx = 3.14159265358979323846 * ((double)(H - 1) - ((double)j + 0.5)) / (double)H;
        const double sinw = sx > 0.0 ? std::sin(sx) : 0.0;
        double rs = 0.0;
        for (size_t i = 0; i < W; ++i) {
            weight[j * W + i] = (weight[j * W + i] + 0.05 * mean) * sinw;
            rs += weight[j * W + i];
        }
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
            c += weight[j * W + i];
        }
        cond[j * (W + 1) + W] = 1.0f;
    }
    marginal[H] = 1.0f;
}

// obj-rs 0.7.0 `load_obj::<TexturedVertex, u32>` (what Mesh::from_file calls, core/mesh.rs:46-48): positions,
// texture coordinates and normals are indexed per face corner as v/vt/vn; every polygon must be a triangle
// with all three indices; identical index triples share one output vertex, created in first-seen order.
bool load_obj_file(const char* path, HostMesh& out, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        err = std::string("cannot open ") + path;
        return false;
    }
    std::vector<float> pos, tex, nrm;
    struct Key {
        int64_t p, t, n;
        bool operator==(const Key& o) const { return p == o.p && t == o.t && n == o.n; }
    };
    struct KeyHash {
        size_t operator()(const Key& k) const {
            return (size_t)(k.p * 73856093) ^ (size_t)(k.t * 19349663) ^ (size_t)(k.n * 83492791);
        }
    };
    std::unordered_map<Key, uint32_t, KeyHash> seen;
    std::vector<Key> order;
    out = HostMesh();
    // the whole file in memory, one NUL-terminated line at a time (lines of any length)
    std::vector<char> text;
    {
        char buf[65536];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.insert(text.end(), buf, buf + got);
        text.push_back('\n');
        text.push_back(0);
    }
    bool ok = true;
    for (size_t line_begin = 0; ok && line_begin + 1 < text.size();) {
        size_t line_end = line_begin;
        while (text[line_end] != '\n' && text[line_end] != 0) ++line_end;
        text[line_end] = 0;
        char* s = text.data() + line_begin;
        line_begin = line_end + 1;
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            float x = 0, y = 0, z = 0;
            if (std::sscanf(s + 1, "%f %f %f", &x, &y, &z) < 3) { err = "malformed v line"; ok = false; break; }
            pos.push_back(x); pos.push_back(y); pos.push_back(z);
        } else if (s[0] == 'v' && s[1] == 't') {
            float u = 0, v = 0;
            if (std::sscanf(s + 2, "%f %f", &u, &v) < 1) { err = "malformed vt line"; ok = false; break; }
            tex.push_back(u); tex.push_back(v);
        } else if (s[0] == 'v' && s[1] == 'n') {
            float x = 0, y = 0, z = 0;
            if (std::sscanf(s + 2, "%f %f %f", &x, &y, &z) < 3) { err = "malformed vn line"; ok = false; break; }
            nrm.push_back(x); nrm.push_back(y); nrm.push_back(z);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            int corners = 0;
            char* p = s + 1;
            while (true) {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == 0 || *p == '\n' || *p == '\r') break;
                long long a = 0, b = 0, c = 0;
                int consumed = 0;
                if (std::sscanf(p, "%lld/%lld/%lld%n", &a, &b, &c, &consumed) < 3) {
                    err = "TexturedVertex needs position/texture/normal on every face corner";
                    ok = false;
                    break;
                }
                p += consumed;
                if (++corners > 3) { err = "model should be triangulated first to be loaded properly"; ok = false; break; }
                // negative indices are relative to the elements read so far
                Key k;
                k.p = a > 0 ? a - 1 : (int64_t)(pos.size() / 3) + a;
                k.t = b > 0 ? b - 1 : (int64_t)(tex.size() / 2) + b;
                k.n = c > 0 ? c - 1 : (int64_t)(nrm.size() / 3) + c;
                if (k.p < 0 || k.t < 0 || k.n < 0) { err = "face index out of range"; ok = false; break; }
                auto it = seen.find(k);
                uint32_t idx;
                if (it == seen.end()) {
                    idx = (uint32_t)order.size();
                    seen.emplace(k, idx);
                    order.push_back(k);
                } else {
                    idx = it->second;
                }
                out.idx.push_back(idx);
            }
            if (ok && corners != 3) { err = "model should be triangulated first to be loaded properly"; ok = false; }
        }
    }
    std::fclose(f);
    if (!ok) return false;
    out.n_vertices = (uint32_t)order.size();
    out.pos.resize(3 * order.size());
    out.uv.resize(2 * order.size());
    out.nrm.resize(3 * order.size());
    for (size_t i = 0; i < order.size(); ++i) {
        const Key& k = order[i];
        if ((uint64_t)k.p >= pos.size() / 3) { err = "face refers to a missing position"; return false; }
        if ((uint64_t)k.t >= tex.size() / 2) { err = "face refers to a missing texture coordinate"; return false; }
        if ((uint64_t)k.n >= nrm.size() / 3) { err = "face refers to a missing normal"; return false; }
        for (int a = 0; a < 3; ++a) out.pos[3 * i + a] = pos[3 * k.p + a];
        for (int a = 0; a < 2; ++a) out.uv[2 * i + a] = tex[2 * k.t + a];
        for (int a = 0; a < 3; ++a) out.nrm[3 * i + a] = nrm[3 * k.n + a];
    }
    return true;
}

}  // namespace vr
