// abi.cu — implementation of include/voidray_cuda.h: scene builder, commit (flatten + upload), the
// progressive wavefront driver behind vr_render_accumulate, resolve, and the gate entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/voidray_cuda.h"
#include "image_io.h"
#include "kernels.cuh"
#include "layout.h"
#include "scene_build.h"

using namespace vr;

namespace {

thread_local std::string g_error;

int32_t fail(int32_t code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define VR_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? VR_ERR_OOM : VR_ERR_CUDA,                     \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                            \
        }                                                                                               \
    } while (0)

struct DeviceBuffers {
    std::vector<void*> ptrs;
    cudaError_t alloc(void** p, size_t bytes) {
        *p = nullptr;
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

}  // namespace

struct vr_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    LaunchDims dims;
    // vr_context_create_multi: the group's other devices (owned). This object is the group's first device.
    std::vector<vr_context*> members;
    std::vector<char> peer_ok;  // per member: the first device can load from it directly (cudaDeviceEnablePeerAccess)
};

// Device allocations of a scene are kept across commits and handed out again in the same order, so a
// re-commit of an unchanged-size scene performs no cudaMalloc / cudaFree.
struct DeviceArena {
    struct Slot {
        void* ptr;
        size_t capacity;
    };
    std::vector<Slot> slots;
    size_t cursor = 0;
    cudaError_t get(void** p, size_t bytes) {
        if (bytes == 0) bytes = 16;
        if (cursor == slots.size()) slots.push_back(Slot{nullptr, 0});
        Slot& s = slots[cursor++];
        if (s.capacity < bytes) {
            if (s.ptr) cudaFree(s.ptr);
            s.ptr = nullptr;
            s.capacity = 0;
            cudaError_t e = cudaMalloc(&s.ptr, bytes);
            if (e != cudaSuccess) return e;
            s.capacity = bytes;
        }
        *p = s.ptr;
        return cudaSuccess;
    }
    void rewind() { cursor = 0; }
    void release() {
        for (Slot& s : slots)
            if (s.ptr) cudaFree(s.ptr);
        slots.clear();
        cursor = 0;
    }
};

struct vr_scene {
    vr_context* ctx = nullptr;
    HostScene host;
    FlatScene flat;
    DeviceArena dev_mem;
    std::vector<const void*> pinned;  // host texture storage registered with cudaHostRegister
    DeviceScene dev;
    std::vector<TextureRec> dev_textures;
    bool committed = false;
    uint64_t commit_serial = 0;
    uint64_t h2d_bytes = 0;
    double flatten_ms = 0.0, upload_ms = 0.0;
    // device group: the same scene on the group's other devices (owned). A replica has no host description of its
    // own: commits upload the primary's (`source`).
    std::vector<vr_scene*> replicas;
    vr_scene* source = nullptr;
};

struct vr_render {
    vr_scene* scene = nullptr;
    uint32_t width = 0, height = 0, n_pixels = 0;
    vr_render_settings settings;
    DeviceBuffers dev_mem;
    float4* accum = nullptr;
    float4* partial = nullptr;
    float4* resolved = nullptr;
    Wavefront wf;
    // Further wavefronts on streams of their own: consecutive batches rotate through them, so that the sparse deep
    // levels and the ramp-down of every kernel of one batch (a persistent grid ends with a few long rays on an almost
    // empty GPU) are filled by the other batch's kernels. Accumulation stays in batch order (ev_acc).
    static constexpr int MAX_WAVEFRONTS = 4;
    Wavefront wf_more[MAX_WAVEFRONTS - 1];                 // wavefronts 1 .. n_wavefronts - 1
    cudaStream_t stream_more[MAX_WAVEFRONTS - 1] = {};     // their streams (wavefront 0 runs on the context's stream)
    cudaEvent_t ev_acc[MAX_WAVEFRONTS] = {};
    int n_wavefronts = 1;
    uint32_t samples_per_batch = 1;
    uint32_t samples_done = 0;
    std::atomic<int> cancel{0};
    std::atomic<int> running{0};  // an accumulate is in flight: only then does vr_render_cancel latch
    // statistics
    std::mutex stats_mutex;
    double seconds = 0.0, device_ms = 0.0, trace_ms = 0.0;
    std::atomic<uint64_t> trace_launches{0}, kernel_launches{0};  // bumped by the render thread, read by vr_render_stats
    unsigned long long segments_host = 0;
    // camera rays that miss the scene's bounds are finished by k_raygen (kernels.cu): on while it pays. A call that
    // culled less than 15 % of its camera rays switches it off; every 16th call after that probes again (the scene or
    // the camera may have changed: the state survives clears and commits)
    unsigned long long culled_host = 0;
    bool cull_camera_rays = true;
    uint32_t calls_without_cull = 0;
    int cull_forced = -1;  // diagnostic: VOIDRAY_CAMERA_CULL=0 / 1 (read by vr_render_begin) pins it off / on
    std::vector<cudaEvent_t> events;  // pairs around trace launches, reused call to call
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // peer accumulation buffers opened through CUDA IPC: (handle bytes, mapped pointer)
    std::vector<std::pair<std::string, void*>> ipc_peers;
    // gate scratch
    uint32_t* dbg_surface = nullptr;
    uint32_t* dbg_prim = nullptr;
    float* dbg_t = nullptr;
    // device group: this render's shards on the group's other devices (owned); `reduced` / `staged` live on the first
    // device (the sum of all shards for read_accum; copies of the shards when a device cannot be peer-mapped)
    std::vector<vr_render*> shards;
    float4* reduced = nullptr;
    std::vector<float4*> staged;
    bool is_shard = false;
};

namespace {

FrameParams frame_params(const vr_render* r) {
    FrameParams fp;
    fp.width = r->width;
    fp.height = r->height;
    fp.pixel_mapping = r->settings.pixel_mapping;
    fp.max_bounces = r->settings.max_bounces;
    fp.firefly_clamp = r->settings.firefly_clamp;
    fp.render_mode = r->settings.render_mode;
    fp.integrator = r->settings.integrator;
    fp.seed = r->settings.seed;
    return fp;
}

// One wavefront batch: ray generation, then per depth closest-hit + shade/compact. No host sync.
void run_wavefront(vr_render* r, const Wavefront& wf, cudaStream_t stream, const PathSource& src, uint32_t n_paths,
                   bool time_trace, size_t* event_cursor) {
    vr_scene* sc = r->scene;
    vr_context* ctx = sc->ctx;
    const FrameParams fp = frame_params(r);
    cudaMemsetAsync(wf.counts, 0, sizeof(uint32_t) * (2 * (fp.max_bounces + 2) + 1), stream);  // counts + cursors + miss_count
    launch_raygen(sc->dev, wf, src, fp, n_paths, r->cull_camera_rays, ctx->dims, stream);
    r->kernel_launches += 1;
    for (uint32_t depth = 0; depth < fp.max_bounces; ++depth) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (time_trace) {
            while (r->events.size() < *event_cursor + 2) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                r->events.push_back(e);
            }
            e0 = r->events[(*event_cursor)++];
            e1 = r->events[(*event_cursor)++];
        }
        if (time_trace) cudaEventRecord(e0, stream);
        launch_trace(sc->dev, wf, depth, n_paths, ctx->dims, stream);
        if (time_trace) cudaEventRecord(e1, stream);
        launch_shade(sc->dev, wf, src, fp, depth, n_paths, ctx->dims, stream);
        r->kernel_launches += 2;
        r->trace_launches += 1;
    }
    if (fp.max_bounces > 1) {  // the misses of depth >= 1 that k_shade parked
        launch_miss(sc->dev, wf, fp, n_paths, ctx->dims, stream);
        r->kernel_launches += 1;
    }
}

// Page-lock the scene's own copy of a texture so that commits DMA straight from it.
void pin_host(vr_scene* scene, const std::vector<float>& v) {
    if (v.empty()) return;
    if (cudaHostRegister((void*)v.data(), v.size() * sizeof(float), cudaHostRegisterPortable) == cudaSuccess)
        scene->pinned.push_back(v.data());
    else
        cudaGetLastError();  // pageable copies still work, just slower
}
void unpin_host(vr_scene* scene, const std::vector<float>& v) {
    for (size_t i = 0; i < scene->pinned.size(); ++i)
        if (scene->pinned[i] == v.data()) {
            cudaHostUnregister((void*)v.data());
            scene->pinned.erase(scene->pinned.begin() + i);
            return;
        }
}

// The RGB f32 texels go over the bus as they are (12 B/texel) and are widened to the 16-byte RGBA
// records the kernels fetch by a device kernel. `stage` is a device scratch buffer of >= 12 * n bytes.
int32_t upload_texture(vr_scene* scene, const HostTexture& t, void* stage, TextureRec* rec) {
    const size_t n = (size_t)t.w * t.h;
    void* d = nullptr;
    if (!t.rgba8.empty()) {  // 8-bit source: 4 B/texel over the bus and in HBM, widened per tap in the kernel
        VR_CUDA(scene->dev_mem.get(&d, 4 * n));
        VR_CUDA(cudaMemcpyAsync(d, t.rgba8.data(), 4 * n, cudaMemcpyHostToDevice, scene->ctx->stream));
        scene->h2d_bytes += 4 * n;
        rec->texels = d;
        rec->width = t.w;
        rec->height = t.h;
        rec->sample_type = t.sample_type;
        rec->pad = 1;
        return VR_OK;
    }
    VR_CUDA(scene->dev_mem.get(&d, 16 * n));
    VR_CUDA(cudaMemcpyAsync(stage, t.rgb.data(), 12 * n, cudaMemcpyHostToDevice, scene->ctx->stream));
    launch_expand_rgb((const float*)stage, (float4*)d, n, scene->ctx->stream);
    scene->h2d_bytes += 12 * n;
    rec->texels = d;
    rec->width = t.w;
    rec->height = t.h;
    rec->sample_type = t.sample_type;
    rec->pad = 0;
    return VR_OK;
}

template <typename V>
int32_t upload_vector(vr_scene* scene, const V& v, const void** out) {
    void* d = nullptr;
    const size_t bytes = v.size() * sizeof(typename V::value_type);
    VR_CUDA(scene->dev_mem.get(&d, bytes));
    if (!v.empty()) VR_CUDA(cudaMemcpyAsync(d, v.data(), bytes, cudaMemcpyHostToDevice, scene->ctx->stream));
    scene->h2d_bytes += bytes;
    *out = d;
    return VR_OK;
}

// Sampling tables of integrator 1, built and uploaded the first time a render asks for them after a commit.
int32_t ensure_env_tables(vr_scene* scene) {
    const HostScene& host = scene->source ? scene->source->host : scene->host;  // a group replica uploads the primary's
    if (host.env_kind != 2 || scene->dev.env_marginal) return VR_OK;
    std::vector<float> marginal, cond;
    build_env_tables(host.env_image, marginal, cond);
    int32_t rc;
    if ((rc = upload_vector(scene, marginal, (const void**)&scene->dev.env_marginal))) return rc;
    if ((rc = upload_vector(scene, cond, (const void**)&scene->dev.env_cond))) return rc;
    VR_CUDA(cudaStreamSynchronize(scene->ctx->stream));  // the host vectors go out of scope
    return VR_OK;
}


template <typename T>
int32_t add_texture_int(vr_scene* scene, const T* pixels, uint32_t w, uint32_t h, uint32_t channels,
                               int32_t sample_type, float denom, uint32_t* texture) {
    if (!pixels || w == 0 || h == 0) return fail(VR_ERR_INVALID, "empty texture");
    if (channels != 3 && channels != 4) return fail(VR_ERR_INVALID, "channels must be 3 or 4");
    std::vector<float> rgb((size_t)3 * w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i)
        for (int c = 0; c < 3; ++c) rgb[3 * i + c] = (float)pixels[channels * i + c] / denom;  // to_rgb32f
    return vr_scene_add_texture_rgb32f(scene, rgb.data(), w, h, sample_type, texture);
}

// No exception leaves the library: every int32_t entry point is a function-try-block ending in VR_CATCH.
int32_t translate_exception() {
    try {
        throw;
    } catch (const std::bad_alloc&) {
        return fail(VR_ERR_OOM, "host allocation failed");
    } catch (const std::exception& e) {
        return fail(VR_ERR_INVALID, std::string("internal error: ") + e.what());
    } catch (...) {
        return fail(VR_ERR_INVALID, "internal error: unknown exception");
    }
}
#define VR_CATCH \
    catch (...) { return translate_exception(); }

int32_t check_scene(vr_scene* s) {
    if (!s) return fail(VR_ERR_INVALID, "null scene");
    return VR_OK;
}

}  // namespace

extern "C" {

const char* vr_last_error(void) { return g_error.c_str(); }
uint32_t vr_abi_version(void) { return 2; }

int32_t vr_context_create(int32_t device, void* cuda_stream, vr_context** out) try {
    if (!out) return fail(VR_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VR_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                     " (this library has no CPU fallback)");
    if (device < 0 || device >= count) return fail(VR_ERR_INVALID, "device index out of range");
    VR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(VR_ERR_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                     std::to_string(prop.minor) + "; this library ships sm_100a code only");
    vr_context* ctx = new (std::nothrow) vr_context();
    if (!ctx) return fail(VR_ERR_OOM, "host allocation failed");
    ctx->device = device;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return fail(VR_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    query_launch_dims(&ctx->dims);
    *out = ctx;
    return VR_OK;
} VR_CATCH

int32_t vr_context_destroy(vr_context* ctx) try {
    if (!ctx) return VR_OK;
    for (vr_context* m : ctx->members) vr_context_destroy(m);
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VR_OK;
} VR_CATCH

// A device group in one process (vr_context_create(device_ids, n_devices)): the returned context is
// the group's first device; scenes created on it are replicated to every device at commit, renders shard their
// samples over the devices and read_accum / resolve sum the shards on the first device over peer memory.
int32_t vr_context_create_multi(const int32_t* device_ids, uint32_t n_devices, vr_context** out) try {
    if (!out) return fail(VR_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!device_ids || n_devices == 0) return fail(VR_ERR_INVALID, "empty device list");
    if (n_devices > (uint32_t)MAX_PEERS + 1) return fail(VR_ERR_INVALID, "too many devices (max 16)");
    vr_context* first = nullptr;
    int32_t rc = vr_context_create(device_ids[0], nullptr, &first);
    if (rc) return rc;
    for (uint32_t i = 1; i < n_devices; ++i) {
        vr_context* m = nullptr;
        rc = vr_context_create(device_ids[i], nullptr, &m);
        if (rc) {
            const std::string why = g_error;
            vr_context_destroy(first);
            return fail(rc, why);
        }
        first->members.push_back(m);
        // peer mapping first device <- member (NVLink / NVSwitch on a B200 box); without it the reduce stages copies
        int can = 0;
        cudaSetDevice(first->device);
        char ok = m->device == first->device ? 1 : 0;  // (a device may be listed twice: two shards on one GPU)
        if (!ok && cudaDeviceCanAccessPeer(&can, first->device, m->device) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->device, 0);
            ok = (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) ? 1 : 0;
        }
        cudaGetLastError();
        first->peer_ok.push_back(ok);
    }
    *out = first;
    return VR_OK;
} VR_CATCH

int32_t vr_context_device_count(vr_context* ctx, uint32_t* n_devices) try {
    if (!ctx || !n_devices) return fail(VR_ERR_INVALID, "null argument");
    *n_devices = 1u + (uint32_t)ctx->members.size();
    return VR_OK;
} VR_CATCH

int32_t vr_scene_create(vr_context* ctx, vr_scene** out) try {
    if (!ctx || !out) return fail(VR_ERR_INVALID, "null argument");
    vr_scene* s = new (std::nothrow) vr_scene();
    if (!s) return fail(VR_ERR_OOM, "host allocation failed");
    s->ctx = ctx;
    // Scene::empty(), core/scene.rs:95-111
    const float eye[3] = {1.0f, 0.0f, 10.0f}, center[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    HostCamera& c = s->host.camera;
    std::memcpy(c.eye, eye, 12);
    camera_look_at(eye, center, up, c.direction, c.up);
    c.fov = 3.14159265358979323846f / 6.0f;
    c.has_dof = 0;
    c.aperture = 0.0f;
    c.focal_point[0] = c.focal_point[1] = c.focal_point[2] = 0.0f;
    for (vr_context* m : ctx->members) {  // device group: one replica per further device
        vr_scene* rep = new (std::nothrow) vr_scene();
        if (!rep) {
            vr_scene_destroy(s);
            return fail(VR_ERR_OOM, "host allocation failed");
        }
        rep->ctx = m;
        rep->source = s;
        s->replicas.push_back(rep);
    }
    *out = s;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_destroy(vr_scene* scene) try {
    if (!scene) return VR_OK;
    for (vr_scene* rep : scene->replicas) vr_scene_destroy(rep);
    cudaSetDevice(scene->ctx->device);
    cudaStreamSynchronize(scene->ctx->stream);
    scene->dev_mem.release();
    for (const void* p : scene->pinned) cudaHostUnregister((void*)p);
    delete scene;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_texture_rgb32f(vr_scene* scene, const float* rgb, uint32_t w, uint32_t h, int32_t sample_type,
                                    uint32_t* texture) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!rgb || w == 0 || h == 0) return fail(VR_ERR_INVALID, "empty texture");
    if ((uint64_t)w * h >= (1ull << 31)) return fail(VR_ERR_INVALID, "texture too large");
    if (sample_type != 0 && sample_type != 1) return fail(VR_ERR_INVALID, "sample_type must be 0 (nearest) or 1 (bilinear)");
    HostTexture t;
    t.w = w;
    t.h = h;
    t.sample_type = sample_type;
    t.rgb.assign(rgb, rgb + (size_t)3 * w * h);
    pack_texture_rgba8(t);
    scene->host.textures.push_back(std::move(t));
    cudaSetDevice(scene->ctx->device);
    pin_host(scene, scene->host.textures.back().rgb);
    {
        const std::vector<uint8_t>& p8 = scene->host.textures.back().rgba8;
        if (!p8.empty() && cudaHostRegister((void*)p8.data(), p8.size(), cudaHostRegisterPortable) == cudaSuccess)
            scene->pinned.push_back(p8.data());
        else
            cudaGetLastError();
    }
    scene->committed = false;
    if (texture) *texture = (uint32_t)scene->host.textures.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_texture_rgb8(vr_scene* scene, const uint8_t* pixels, uint32_t w, uint32_t h, uint32_t channels,
                                  int32_t sample_type, uint32_t* texture) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    return add_texture_int(scene, pixels, w, h, channels, sample_type, 255.0f, texture);
} VR_CATCH
int32_t vr_scene_add_texture_rgb16(vr_scene* scene, const uint16_t* pixels, uint32_t w, uint32_t h, uint32_t channels,
                                   int32_t sample_type, uint32_t* texture) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    return add_texture_int(scene, pixels, w, h, channels, sample_type, 65535.0f, texture);
} VR_CATCH

int32_t vr_image_load_rgb32f(const char* path, uint32_t* w, uint32_t* h, float** rgb) try {
    if (!path || !w || !h || !rgb) return fail(VR_ERR_INVALID, "null argument");
    *rgb = nullptr;
    DecodedImage img;
    std::string err;
    if (!decode_image_file(path, img, err)) return fail(VR_ERR_INVALID, std::string(path) + ": " + err);
    std::vector<float> f;
    image_to_rgb32f(img, f);
    float* out = (float*)std::malloc(f.size() * sizeof(float));
    if (!out) return fail(VR_ERR_OOM, "host allocation failed");
    std::memcpy(out, f.data(), f.size() * sizeof(float));
    *w = img.w;
    *h = img.h;
    *rgb = out;
    return VR_OK;
} VR_CATCH

int32_t vr_image_free(float* rgb) try {
    std::free(rgb);
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_image_texture_file(vr_scene* scene, const char* path, int32_t sample_type, uint32_t* texture) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!path) return fail(VR_ERR_INVALID, "null path");
    DecodedImage img;
    std::string err;
    if (!decode_image_file(path, img, err)) return fail(VR_ERR_INVALID, std::string(path) + ": " + err);
    std::vector<float> f;
    image_to_rgb32f(img, f);
    return vr_scene_add_texture_rgb32f(scene, f.data(), img.w, img.h, sample_type, texture);
} VR_CATCH

int32_t vr_scene_add_mesh(vr_scene* scene, const float* positions, const float* uvs, const float* normals,
                          uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices, uint32_t* surface) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if ((!positions && n_vertices) || (!indices && n_indices)) return fail(VR_ERR_INVALID, "null mesh buffers");
    const uint32_t n_idx = n_indices - n_indices % 3;  // chunks_exact(3), mesh.rs:79
    for (uint32_t i = 0; i < n_idx; ++i)
        if (indices[i] >= n_vertices) return fail(VR_ERR_INVALID, "mesh index out of range (the reference would panic)");
    HostMesh m;
    m.n_vertices = n_vertices;
    m.pos.assign(positions, positions + (size_t)3 * n_vertices);
    if (uvs) m.uv.assign(uvs, uvs + (size_t)2 * n_vertices);
    else m.uv.assign((size_t)2 * n_vertices, 0.0f);
    if (normals) m.nrm.assign(normals, normals + (size_t)3 * n_vertices);
    else m.nrm.assign((size_t)3 * n_vertices, 0.0f);
    m.idx.assign(indices, indices + n_idx);
    scene->host.meshes.push_back(std::move(m));
    HostSurface sf;
    sf.kind = 0;
    sf.mesh = (uint32_t)scene->host.meshes.size() - 1;
    scene->host.surfaces.push_back(sf);
    scene->committed = false;
    if (surface) *surface = (uint32_t)scene->host.surfaces.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_mesh_from_obj_file(vr_scene* scene, const char* path, uint32_t* surface, uint32_t* n_vertices,
                                        uint32_t* n_triangles) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!path) return fail(VR_ERR_INVALID, "null path");
    HostMesh m;
    std::string err;
    if (!load_obj_file(path, m, err)) return fail(VR_ERR_INVALID, std::string(path) + ": " + err);
    if (n_vertices) *n_vertices = m.n_vertices;
    if (n_triangles) *n_triangles = (uint32_t)(m.idx.size() / 3);
    scene->host.meshes.push_back(std::move(m));
    HostSurface sf;
    sf.kind = 0;
    sf.mesh = (uint32_t)scene->host.meshes.size() - 1;
    scene->host.surfaces.push_back(sf);
    scene->committed = false;
    if (surface) *surface = (uint32_t)scene->host.surfaces.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_obj_load(const char* path, vr_obj_mesh* out) try {
    if (!path || !out) return fail(VR_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof *out);
    HostMesh m;
    std::string err;
    if (!load_obj_file(path, m, err)) return fail(VR_ERR_INVALID, std::string(path) + ": " + err);
    auto dup = [](const void* src, size_t bytes) -> void* {
        void* p = std::malloc(bytes ? bytes : 1);
        if (p && bytes) std::memcpy(p, src, bytes);
        return p;
    };
    out->n_vertices = m.n_vertices;
    out->n_indices = (uint32_t)m.idx.size();
    out->positions = (float*)dup(m.pos.data(), m.pos.size() * sizeof(float));
    out->uvs = (float*)dup(m.uv.data(), m.uv.size() * sizeof(float));
    out->normals = (float*)dup(m.nrm.data(), m.nrm.size() * sizeof(float));
    out->indices = (uint32_t*)dup(m.idx.data(), m.idx.size() * sizeof(uint32_t));
    if (!out->positions || !out->uvs || !out->normals || !out->indices) {
        vr_obj_free(out);
        return fail(VR_ERR_OOM, "host allocation failed");
    }
    return VR_OK;
} VR_CATCH

int32_t vr_obj_free(vr_obj_mesh* mesh) try {
    if (!mesh) return VR_OK;
    std::free(mesh->positions);
    std::free(mesh->uvs);
    std::free(mesh->normals);
    std::free(mesh->indices);
    std::memset(mesh, 0, sizeof *mesh);
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_sphere(vr_scene* scene, const float center[3], float radius, uint32_t* surface) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!center) return fail(VR_ERR_INVALID, "null center");
    HostSurface sf;
    sf.kind = 1;
    std::memcpy(sf.center, center, 12);
    sf.radius_or_height = radius;
    scene->host.surfaces.push_back(sf);
    scene->committed = false;
    if (surface) *surface = (uint32_t)scene->host.surfaces.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_ground_plane(vr_scene* scene, float height, uint32_t* surface) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    HostSurface sf;
    sf.kind = 2;
    sf.radius_or_height = height;
    scene->host.surfaces.push_back(sf);
    scene->committed = false;
    if (surface) *surface = (uint32_t)scene->host.surfaces.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_material(vr_scene* scene, const vr_material_desc* desc, uint32_t* material) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!desc) return fail(VR_ERR_INVALID, "null material desc");
    if (desc->kind < 0 || desc->kind > VR_MAT_MICROFACET) return fail(VR_ERR_INVALID, "unknown material kind");
    MaterialRec m;
    m.kind = desc->kind;
    std::memcpy(m.color, desc->color, 12);
    m.param = desc->param;
    m.albedo_tex = desc->kind == VR_MAT_LAMBERTIAN ? desc->albedo_tex : -1;
    m.normal_tex = desc->kind == VR_MAT_LAMBERTIAN ? desc->normal_tex : -1;
    m.transparent = desc->transparent ? 1 : 0;
    m.index = desc->index;
    m.roughness = desc->roughness;
    m.metallic = desc->metallic;
    m.emittance = desc->emittance;
    if (desc->kind == VR_MAT_EMISSION) {  // Emission::new: color * strength, simple.rs:168-172
        m.color[0] = desc->color[0] * desc->param;
        m.color[1] = desc->color[1] * desc->param;
        m.color[2] = desc->color[2] * desc->param;
    }
    scene->host.materials.push_back(m);
    scene->committed = false;
    if (material) *material = (uint32_t)scene->host.materials.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_add_object(vr_scene* scene, uint32_t material, uint32_t surface, uint32_t* object) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (material >= scene->host.materials.size()) return fail(VR_ERR_INVALID, "unknown material handle");
    if (surface >= scene->host.surfaces.size()) return fail(VR_ERR_INVALID, "unknown surface handle");
    scene->host.objects.push_back(HostObject{surface, material});
    scene->committed = false;
    if (object) *object = (uint32_t)scene->host.objects.size() - 1;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_set_camera(vr_scene* scene, const float eye[3], const float direction[3], const float up[3],
                            float fov, int32_t has_dof, float aperture, const float focal_point[3]) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!eye || !direction || !up) return fail(VR_ERR_INVALID, "null camera vector");
    if (has_dof && !focal_point) return fail(VR_ERR_INVALID, "has_dof needs a focal point");
    HostCamera& c = scene->host.camera;
    std::memcpy(c.eye, eye, 12);
    std::memcpy(c.direction, direction, 12);
    std::memcpy(c.up, up, 12);
    c.fov = fov;
    c.has_dof = has_dof ? 1 : 0;
    c.aperture = aperture;
    if (focal_point) std::memcpy(c.focal_point, focal_point, 12);
    scene->committed = false;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_set_camera_look_at(vr_scene* scene, const float eye[3], const float center[3], const float up[3],
                                    float fov) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!eye || !center || !up) return fail(VR_ERR_INVALID, "null camera vector");
    HostCamera& c = scene->host.camera;
    std::memcpy(c.eye, eye, 12);
    camera_look_at(eye, center, up, c.direction, c.up);
    c.fov = fov;
    c.has_dof = 0;
    scene->committed = false;
    return VR_OK;
} VR_CATCH

int32_t vr_camera_look_at(const float eye[3], const float center[3], const float up[3], float direction_out[3],
                          float up_out[3]) try {
    if (!eye || !center || !up || !direction_out || !up_out) return fail(VR_ERR_INVALID, "null camera vector");
    camera_look_at(eye, center, up, direction_out, up_out);
    return VR_OK;
} VR_CATCH

int32_t vr_scene_set_environment_uniform(vr_scene* scene, const float rgb[3]) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!rgb) return fail(VR_ERR_INVALID, "null colour");
    scene->host.env_kind = 1;
    std::memcpy(scene->host.env_color, rgb, 12);
    unpin_host(scene, scene->host.env_image.rgb);
    scene->host.env_image = HostTexture();
    scene->committed = false;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_set_environment_hdri_rgb32f(vr_scene* scene, const float* rgb, uint32_t w, uint32_t h) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!rgb || w == 0 || h == 0) return fail(VR_ERR_INVALID, "empty environment image");
    if ((uint64_t)w * h >= (1ull << 31)) return fail(VR_ERR_INVALID, "environment image too large");
    scene->host.env_kind = 2;
    cudaSetDevice(scene->ctx->device);
    unpin_host(scene, scene->host.env_image.rgb);
    scene->host.env_image.w = w;
    scene->host.env_image.h = h;
    scene->host.env_image.sample_type = 1;
    scene->host.env_image.rgb.assign(rgb, rgb + (size_t)3 * w * h);
    pin_host(scene, scene->host.env_image.rgb);
    scene->committed = false;
    return VR_OK;
} VR_CATCH

int32_t vr_scene_set_environment_hdri_file(vr_scene* scene, const char* path) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!path) return fail(VR_ERR_INVALID, "null path");
    DecodedImage img;
    std::string err;
    if (!decode_image_file(path, img, err)) return fail(VR_ERR_INVALID, std::string(path) + ": " + err);
    std::vector<float> f;
    image_to_rgb32f(img, f);
    return vr_scene_set_environment_hdri_rgb32f(scene, f.data(), img.w, img.h);
} VR_CATCH

int32_t vr_scene_clear_environment(vr_scene* scene) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    scene->host.env_kind = 0;
    unpin_host(scene, scene->host.env_image.rgb);
    scene->host.env_image = HostTexture();
    scene->committed = false;
    return VR_OK;
} VR_CATCH

// The device half of a commit, in two steps so that the large copies overlap the host's BVH build: the textures and
// the environment do not depend on the flatten, so their copies (from page-locked host memory) are put on the stream
// first and run while the host flattens; the geometry follows. dst == src for a single device; a group's replicas upload
// the primary's host data.
static int32_t upload_textures(vr_scene* dst, const vr_scene* src) {
    VR_CUDA(cudaSetDevice(dst->ctx->device));
    dst->h2d_bytes = 0;
    VR_CUDA(cudaStreamSynchronize(dst->ctx->stream));
    dst->dev_mem.rewind();
    dst->dev_textures.clear();
    DeviceScene& d = dst->dev;
    std::memset(&d, 0, sizeof(d));
    const HostScene& host = src->host;
    int32_t rc;
    size_t max_texels = host.env_kind == 2 ? (size_t)host.env_image.w * host.env_image.h : 0;
    for (const HostTexture& t : host.textures) max_texels = std::max(max_texels, (size_t)t.w * t.h);
    void* stage = nullptr;
    VR_CUDA(dst->dev_mem.get(&stage, 12 * max_texels));
    for (const HostTexture& t : host.textures) {
        TextureRec rec;
        if ((rc = upload_texture(dst, t, stage, &rec))) return rc;
        dst->dev_textures.push_back(rec);
    }
    if ((rc = upload_vector(dst, dst->dev_textures, (const void**)&d.textures))) return rc;
    d.env_kind = host.env_kind;
    std::memcpy(d.env_color, host.env_color, 12);
    if (host.env_kind == 2) {
        if ((rc = upload_texture(dst, host.env_image, stage, &d.env_tex))) return rc;
        // the sampling tables of integrator 1 are built on demand (ensure_env_tables)
    }
    return VR_OK;
}

static int32_t upload_geometry(vr_scene* dst, const vr_scene* src) {
    VR_CUDA(cudaSetDevice(dst->ctx->device));
    DeviceScene& d = dst->dev;
    const FlatScene& f = src->flat;
    const HostScene& host = src->host;
    int32_t rc;
    if ((rc = upload_vector(dst, f.nodes, &d.nodes))) return rc;
    if ((rc = upload_vector(dst, f.tri_isect, &d.tri_isect))) return rc;
    if ((rc = upload_vector(dst, f.tri_shade, &d.tri_shade))) return rc;
    if ((rc = upload_vector(dst, f.tri_surface, (const void**)&d.tri_surface))) return rc;
    if ((rc = upload_vector(dst, f.tri_prim, (const void**)&d.tri_prim))) return rc;
    if ((rc = upload_vector(dst, host.materials, (const void**)&d.materials))) return rc;
    if ((rc = upload_vector(dst, f.analytics, (const void**)&d.analytics))) return rc;
    if ((rc = upload_vector(dst, f.scene_tree, (const void**)&d.scene_tree))) return rc;
    if ((rc = upload_vector(dst, f.surface_node, (const void**)&d.surface_node))) return rc;
    d.n_scene_nodes = (uint32_t)f.scene_tree.size();
    d.n_surfaces = (uint32_t)f.surface_node.size();
    d.has_microfacet = 0;
    for (const MaterialRec& m : host.materials)
        if (m.kind == VR_MAT_MICROFACET) d.has_microfacet = 1;
    std::memcpy(d.grid_min, f.grid_min, 12);
    std::memcpy(d.grid_extent, f.grid_extent, 12);
    d.n_tris = f.n_tris;
    d.n_analytics = (uint32_t)f.analytics.size();
    d.camera = f.camera;
    VR_CUDA(cudaStreamSynchronize(dst->ctx->stream));
    return VR_OK;
}

int32_t vr_scene_commit(vr_scene* scene) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (scene->source) return fail(VR_ERR_INVALID, "a group replica is committed through its primary scene");
    scene->committed = false;  // a commit that fails half-way leaves no usable scene behind
    for (vr_scene* rep : scene->replicas) rep->committed = false;
    const size_t n = 1 + scene->replicas.size();
    auto device_scene = [&](size_t i) { return i == 0 ? scene : scene->replicas[i - 1]; };
    const auto t_begin = std::chrono::steady_clock::now();
    // textures / environment: asynchronous copies from page-locked memory, issued before the host work they overlap
    for (size_t i = 0; i < n; ++i) {
        const int32_t rc = upload_textures(device_scene(i), scene);
        if (rc) return rc;
    }
    const auto t_issued = std::chrono::steady_clock::now();
    std::string err;
    if (!flatten_scene(scene->host, scene->flat, err)) return fail(VR_ERR_INVALID, err);
    const auto t_flat = std::chrono::steady_clock::now();
    if (scene->flat.bvh_depth > 32) return fail(VR_ERR_INVALID, "BVH too deep for the traversal stack");
    if (n == 1) {
        const int32_t rc = upload_geometry(scene, scene);
        if (rc) return rc;
    } else {
        // one host thread per device: the copies of the pageable flat arrays block their thread, the devices' DMA
        // engines run side by side
        std::vector<int32_t> rcs(n, VR_OK);
        std::vector<std::string> errs(n);
        std::vector<std::thread> th;
        auto work = [&](size_t i) {
            try {
                rcs[i] = upload_geometry(device_scene(i), scene);
            } catch (...) {
                rcs[i] = translate_exception();
            }
            if (rcs[i]) errs[i] = g_error;  // thread-local on the worker
        };
        for (size_t i = 1; i < n; ++i) th.emplace_back(work, i);
        work(0);
        for (std::thread& t : th) t.join();
        for (size_t i = 0; i < n; ++i)
            if (rcs[i]) return fail(rcs[i], "device " + std::to_string(i) + " of the group: " + errs[i]);
    }
    const auto t_end = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    scene->flatten_ms = ms(t_issued, t_flat);
    scene->upload_ms = ms(t_begin, t_issued) + ms(t_flat, t_end);  // what the copies add on top of the host build
    scene->committed = true;
    scene->commit_serial++;
    for (vr_scene* rep : scene->replicas) {
        rep->committed = true;
        rep->commit_serial++;
        scene->h2d_bytes += rep->h2d_bytes;
    }
    return VR_OK;
} VR_CATCH

int32_t vr_scene_get_info(vr_scene* scene, vr_scene_info* out) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!out) return fail(VR_ERR_INVALID, "null argument");
    if (!scene->committed) return fail(VR_ERR_INVALID, "scene not committed");
    out->n_triangles = scene->flat.n_tris;
    out->n_bvh_nodes = (uint32_t)(scene->flat.nodes.size() / DEVICE_NODE_QUADS);
    out->bvh_depth = scene->flat.bvh_depth;
    out->n_analytic_surfaces = (uint32_t)scene->flat.analytics.size();
    out->n_textures = (uint32_t)scene->host.textures.size();
    out->reserved = 0;
    out->h2d_bytes = scene->h2d_bytes;
    out->flatten_ms = scene->flatten_ms;
    out->upload_ms = scene->upload_ms;
    return VR_OK;
} VR_CATCH

static int32_t render_begin_single(vr_scene* scene, uint32_t width, uint32_t height, const vr_render_settings* settings,
                                   vr_render** out) {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!settings || !out) return fail(VR_ERR_INVALID, "null argument");
    if (!scene->committed) return fail(VR_ERR_INVALID, "vr_scene_commit must be called before vr_render_begin");
    if (width == 0 || height == 0 || (uint64_t)width * height >= (1ull << 31))
        return fail(VR_ERR_INVALID, "bad target dimensions");
    if (settings->total_samples == 0) return fail(VR_ERR_INVALID, "total_samples must be > 0");
    if (settings->max_bounces > 64) return fail(VR_ERR_INVALID, "max_bounces > 64 is not supported");
    if (settings->integrator != 0 && settings->integrator != 1) return fail(VR_ERR_INVALID, "unknown integrator");
    VR_CUDA(cudaSetDevice(scene->ctx->device));
    if (settings->integrator == 1) {
        const int32_t rc = ensure_env_tables(scene);
        if (rc) return rc;
    }
    vr_render* r = new (std::nothrow) vr_render();
    if (!r) return fail(VR_ERR_OOM, "host allocation failed");
    r->scene = scene;
    r->width = width;
    r->height = height;
    r->n_pixels = width * height;
    r->settings = *settings;

    // Paths in flight per wavefront batch. More is better for the deep, sparse depths (measured: +27 % on
    // config 1, +7 % on config 2 going from 8 Mi to 32 Mi); 32 Mi slots are 7.8 GB of the 180 GB of HBM at 8 bounces.
    // Two wavefronts (see vr_render::wf_more) share max_paths_in_flight when the caller sets it; VOIDRAY_STREAMS=n
    // (experiment knob, 1 .. 4) sets their number.
    const char* streams_env = std::getenv("VOIDRAY_STREAMS");
    int n_wf = streams_env ? std::atoi(streams_env) : 2;
    n_wf = std::max(1, std::min(n_wf, (int)vr_render::MAX_WAVEFRONTS));
    n_wf = (int)std::min<uint64_t>((uint64_t)n_wf, settings->total_samples);
    uint64_t capacity = settings->max_paths_in_flight ? settings->max_paths_in_flight : (32ull << 20);
    if (settings->max_paths_in_flight) {
        while (n_wf > 1 && capacity / n_wf < r->n_pixels) --n_wf;
        capacity /= n_wf;
    }
    const uint64_t samples_each = (settings->total_samples + n_wf - 1) / n_wf;
    capacity = std::min<uint64_t>(capacity, (uint64_t)r->n_pixels * samples_each);
    if (capacity < r->n_pixels) capacity = r->n_pixels;
    r->samples_per_batch = (uint32_t)(capacity / r->n_pixels);
    capacity = (uint64_t)r->samples_per_batch * r->n_pixels;
    if (capacity >= (1ull << 26)) {  // a miss record packs slot | depth << 26 (kernels.cu)
        delete r;
        return fail(VR_ERR_INVALID, "max_paths_in_flight too large");
    }
    r->n_wavefronts = n_wf;
    if (const char* v = std::getenv("VOIDRAY_CAMERA_CULL")) r->cull_forced = std::atoi(v) != 0 ? 1 : 0;
    const size_t cap = capacity;
    const uint32_t levels = settings->max_bounces ? settings->max_bounces : 1;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = r->dev_mem.alloc(p, bytes);
    };
    A((void**)&r->accum, 16ull * r->n_pixels);
    A((void**)&r->partial, 16ull * r->n_pixels);
    A((void**)&r->resolved, 16ull * r->n_pixels);
    unsigned long long* segments = nullptr;
    A((void**)&segments, 16);  // + the culled camera rays
    auto alloc_wavefront = [&](Wavefront& w) {
        w.capacity = (uint32_t)capacity;
        A((void**)&w.ray_o[0], 16 * cap);
        A((void**)&w.ray_o[1], 16 * cap);
        A((void**)&w.ray_d[0], 16 * cap);
        A((void**)&w.ray_d[1], 16 * cap);
        A((void**)&w.hit, 16 * cap);
        A((void**)&w.radiance, 16 * cap);
        A((void**)&w.att, 16 * cap * levels);
        A((void**)&w.queue[0], 4 * cap);
        A((void**)&w.queue[1], 4 * cap);
        A((void**)&w.miss, 16 * cap);
        A((void**)&w.counts, 4 * (2 * (settings->max_bounces + 2) + 1));
        w.segments = segments;  // one counter for the whole render
        w.culled = segments ? segments + 1 : nullptr;
        w.cursors = w.counts ? w.counts + (settings->max_bounces + 2) : nullptr;
        w.miss_count = w.counts ? w.counts + 2 * (settings->max_bounces + 2) : nullptr;
    };
    Wavefront& wf = r->wf;
    alloc_wavefront(wf);
    for (int k = 1; k < n_wf; ++k) {
        alloc_wavefront(r->wf_more[k - 1]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream_more[k - 1], cudaStreamNonBlocking);
    }
    for (int k = 0; k < n_wf && n_wf > 1; ++k)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_acc[k], cudaEventDisableTiming);
    A((void**)&r->dbg_surface, 4ull * r->n_pixels);
    A((void**)&r->dbg_prim, 4ull * r->n_pixels);
    A((void**)&r->dbg_t, 4ull * r->n_pixels);
    if (e == cudaSuccess) e = cudaEventCreate(&r->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&r->ev_end);
    cudaStream_t st = scene->ctx->stream;
    if (e == cudaSuccess) e = cudaMemsetAsync(r->accum, 0, 16ull * r->n_pixels, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(r->partial, 0, 16ull * r->n_pixels, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(wf.segments, 0, 16, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        r->dev_mem.release();
        for (cudaStream_t sm : r->stream_more)
            if (sm) cudaStreamDestroy(sm);
        for (cudaEvent_t ev : r->ev_acc)
            if (ev) cudaEventDestroy(ev);
        delete r;
        return fail(e == cudaErrorMemoryAllocation ? VR_ERR_OOM : VR_ERR_CUDA,
                    std::string("vr_render_begin: ") + cudaGetErrorString(e));
    }
    *out = r;
    return VR_OK;
}

int32_t vr_render_begin(vr_scene* scene, uint32_t width, uint32_t height, const vr_render_settings* settings,
                        vr_render** out) try {
    if (!out) return fail(VR_ERR_INVALID, "null argument");
    *out = nullptr;
    if (scene && scene->source) return fail(VR_ERR_INVALID, "a group replica renders through its primary scene");
    vr_render* r = nullptr;
    int32_t rc = render_begin_single(scene, width, height, settings, &r);
    if (rc) return rc;
    // device group: one shard per further device (its own wavefront state and accumulation buffer), plus the buffer
    // on the first device that receives the sum for vr_render_read_accum
    for (size_t i = 0; i < scene->replicas.size(); ++i) {
        vr_render* sh = nullptr;
        rc = render_begin_single(scene->replicas[i], width, height, settings, &sh);
        if (rc) {
            const std::string why = g_error;
            vr_render_end(r);
            return fail(rc, why);
        }
        sh->is_shard = true;
        r->shards.push_back(sh);
    }
    if (!r->shards.empty()) {
        cudaSetDevice(scene->ctx->device);
        cudaError_t e = r->dev_mem.alloc((void**)&r->reduced, 16ull * r->n_pixels);
        for (size_t i = 0; i < r->shards.size() && e == cudaSuccess; ++i) {
            float4* st = nullptr;
            if (!scene->ctx->peer_ok[i]) e = r->dev_mem.alloc((void**)&st, 16ull * r->n_pixels);
            r->staged.push_back(st);
        }
        if (e != cudaSuccess) {
            vr_render_end(r);
            return fail(e == cudaErrorMemoryAllocation ? VR_ERR_OOM : VR_ERR_CUDA, std::string("vr_render_begin: ") + cudaGetErrorString(e));
        }
    }
    *out = r;
    return VR_OK;
} VR_CATCH

int32_t vr_render_end(vr_render* r) try {
    if (!r) return VR_OK;
    for (vr_render* sh : r->shards) vr_render_end(sh);
    cudaSetDevice(r->scene->ctx->device);
    cudaStreamSynchronize(r->scene->ctx->stream);
    for (auto& p : r->ipc_peers) cudaIpcCloseMemHandle(p.second);
    for (cudaEvent_t e : r->events) cudaEventDestroy(e);
    if (r->ev_begin) cudaEventDestroy(r->ev_begin);
    if (r->ev_end) cudaEventDestroy(r->ev_end);
    for (cudaStream_t sm : r->stream_more)
        if (sm) {
            cudaStreamSynchronize(sm);
            cudaStreamDestroy(sm);
        }
    for (cudaEvent_t ev : r->ev_acc)
        if (ev) cudaEventDestroy(ev);
    r->dev_mem.release();
    delete r;
    return VR_OK;
} VR_CATCH

int32_t vr_render_clear(vr_render* r) try {
    if (!r) return fail(VR_ERR_INVALID, "null render");
    for (vr_render* sh : r->shards) {
        const int32_t rc = vr_render_clear(sh);
        if (rc) return rc;
    }
    cudaStream_t st = r->scene->ctx->stream;
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    VR_CUDA(cudaMemsetAsync(r->accum, 0, 16ull * r->n_pixels, st));
    VR_CUDA(cudaMemsetAsync(r->partial, 0, 16ull * r->n_pixels, st));
    VR_CUDA(cudaMemsetAsync(r->wf.segments, 0, 16, st));
    VR_CUDA(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lock(r->stats_mutex);
    r->samples_done = 0;
    r->seconds = r->device_ms = r->trace_ms = 0.0;
    r->trace_launches = r->kernel_launches = 0;
    r->segments_host = 0;
    r->culled_host = 0;
    r->cancel = 0;
    return VR_OK;
} VR_CATCH

// `samples` camera samples per pixel starting at global sample index `first_sample`, on r's own device. Blocking.
// alpha_inc: what the call adds to the alpha channel (1 per iterative_render call; 0 on a group's further shards).
static int32_t accumulate_range(vr_render* r, uint32_t first_sample, uint32_t samples, float alpha_inc, uint32_t* done_out) {
    *done_out = 0;
    vr_context* ctx = r->scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    if (r->settings.integrator == 1) {
        // a re-commit rewinds the scene's device memory and with it the sampling tables of integrator 1
        const int32_t rc = ensure_env_tables(r->scene);
        if (rc) return rc;
    }
    if (r->cull_forced >= 0) r->cull_camera_rays = r->cull_forced != 0;
    else if (!r->cull_camera_rays && ++r->calls_without_cull >= 16u) {
        r->cull_camera_rays = true;
        r->calls_without_cull = 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    const float inv_total = 1.0f / (float)r->settings.total_samples;  // iterative.rs:45
    VR_CUDA(cudaEventRecord(r->ev_begin, ctx->stream));
    const int n_wf = r->n_wavefronts;
    for (int k = 1; k < n_wf; ++k) VR_CUDA(cudaStreamWaitEvent(r->stream_more[k - 1], r->ev_begin, 0));  // after whatever the caller queued before
    uint32_t done = 0;
    size_t event_cursor = 0;
    bool cancelled = false;
    // with n wavefronts a call of >= n samples is cut into at least n batches, so that there is something to overlap
    const uint32_t per_batch = std::min(r->samples_per_batch, std::max(1u, (samples + (uint32_t)n_wf - 1u) / (uint32_t)n_wf));
    int last = -1;  // the wavefront of the previous batch
    for (uint32_t batch = 0; done < samples; ++batch) {
        if (r->cancel.load()) {
            cancelled = true;
            break;
        }
        const int w = (int)(batch % (uint32_t)n_wf);
        const Wavefront& wf = w ? r->wf_more[w - 1] : r->wf;
        cudaStream_t stream = w ? r->stream_more[w - 1] : ctx->stream;
        const uint32_t nb = std::min(samples - done, per_batch);
        const PathSource src = make_path_source(nullptr, nullptr, r->width, r->height, first_sample + done);
        run_wavefront(r, wf, stream, src, nb * r->n_pixels, true, &event_cursor);
        done += nb;
        // the per-pixel sums run in batch order whatever the streams do: bit-identical to a single wavefront
        if (n_wf > 1 && last >= 0) VR_CUDA(cudaStreamWaitEvent(stream, r->ev_acc[last], 0));
        launch_accumulate(wf, r->partial, r->accum, r->width, r->height, nb, done == samples ? 1 : 0, inv_total, alpha_inc,
                          stream);
        if (n_wf > 1) VR_CUDA(cudaEventRecord(r->ev_acc[w], stream));
        last = w;
        r->kernel_launches += 1;
    }
    if (n_wf > 1 && last >= 0) VR_CUDA(cudaStreamWaitEvent(ctx->stream, r->ev_acc[last], 0));
    if (cancelled && done > 0) {
        launch_accumulate(r->wf, r->partial, r->accum, r->width, r->height, 0, 1, inv_total, alpha_inc, ctx->stream);
        r->kernel_launches += 1;
    }
    VR_CUDA(cudaEventRecord(r->ev_end, ctx->stream));
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    VR_CUDA(cudaGetLastError());
    float ms = 0.0f;
    VR_CUDA(cudaEventElapsedTime(&ms, r->ev_begin, r->ev_end));
    // time during which a closest-hit kernel was running: the union of the launches' [start, end] intervals (with two
    // wavefronts the launches of the two streams overlap each other and the other stream's shading)
    double trace_ms = 0.0;
    {
        std::vector<std::pair<float, float>> spans;
        for (size_t i = 0; i + 1 < event_cursor; i += 2) {
            float a = 0.0f, b = 0.0f;
            cudaEventElapsedTime(&a, r->ev_begin, r->events[i]);
            cudaEventElapsedTime(&b, r->ev_begin, r->events[i + 1]);
            spans.emplace_back(a, b);
        }
        std::sort(spans.begin(), spans.end());
        float lo = 0.0f, hi = -1.0f;
        for (const auto& sp : spans) {
            if (hi < lo || sp.first > hi) {
                if (hi >= lo) trace_ms += hi - lo;
                lo = sp.first;
                hi = sp.second;
            } else if (sp.second > hi) {
                hi = sp.second;
            }
        }
        if (hi >= lo) trace_ms += hi - lo;
    }
    unsigned long long counters[2] = {0, 0};  // segments, culled camera rays (both since the last clear)
    VR_CUDA(cudaMemcpy(counters, r->wf.segments, 16, cudaMemcpyDeviceToHost));
    const unsigned long long seg = counters[0];
    if (r->cull_forced < 0 && r->cull_camera_rays && done > 0) {
        const double culled_now = (double)(counters[1] - r->culled_host), camera_rays = (double)r->n_pixels * done;
        if (culled_now < 0.15 * camera_rays) r->cull_camera_rays = false;
    }
    r->culled_host = counters[1];
    const auto t1 = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lock(r->stats_mutex);
        r->device_ms += ms;
        r->trace_ms += trace_ms;
        r->segments_host = seg;
        r->seconds += std::chrono::duration<double>(t1 - t0).count();
    }
    *done_out = done;
    if (cancelled) return fail(VR_ERR_CANCELLED, "accumulate cancelled");
    return VR_OK;
}

int32_t vr_render_accumulate(vr_render* r, uint32_t samples) try {
    if (!r) return fail(VR_ERR_INVALID, "null render");
    if (r->is_shard) return fail(VR_ERR_INVALID, "a group shard is driven through its primary render");
    if (!r->scene->committed) return fail(VR_ERR_INVALID, "scene was edited after commit");
    if (samples == 0) return VR_OK;
    // Cancel acts on the accumulate that is running when it is issued (renderer.rs:101-106 polls between batches of
    // the render in flight): a cancel that arrives while nothing runs is dropped, one that arrives after the last
    // batch check is forgotten when the next accumulate starts.
    r->cancel = 0;
    for (vr_render* sh : r->shards) sh->cancel = 0;
    r->running = 1;
    struct Running {
        std::atomic<int>& flag;
        ~Running() { flag = 0; }
    } running_guard{r->running};
    const uint32_t first = r->settings.sample_offset + r->samples_done;
    uint32_t done = 0;
    int32_t rc = VR_OK;
    if (r->shards.empty()) {
        rc = accumulate_range(r, first, samples, 1.0f, &done);
    } else {
        // device group: the call's sample range is cut into contiguous pieces, one per device, each driven by its own
        // host thread; the random streams are keyed by (pixel, global sample index), so the union of the shards is the
        // sample set one device would have drawn. Only the first device counts the call in the alpha channel.
        const uint32_t n = 1u + (uint32_t)r->shards.size();
        std::vector<uint32_t> off(n), cnt(n), dn(n, 0);
        for (uint32_t g = 0, o = 0; g < n; ++g) {
            cnt[g] = samples / n + (g < samples % n ? 1u : 0u);
            off[g] = o;
            o += cnt[g];
        }
        std::vector<int32_t> rcs(n, VR_OK);
        std::vector<std::string> errs(n);
        auto work = [&](uint32_t g) {
            vr_render* sh = g == 0 ? r : r->shards[g - 1];
            if (cnt[g] == 0) return;
            try {
                rcs[g] = accumulate_range(sh, first + off[g], cnt[g], g == 0 ? 1.0f : 0.0f, &dn[g]);
            } catch (...) {
                rcs[g] = translate_exception();
            }
            if (rcs[g]) errs[g] = g_error;
        };
        std::vector<std::thread> th;
        for (uint32_t g = 1; g < n; ++g) th.emplace_back(work, g);
        work(0);
        for (std::thread& t : th) t.join();
        for (uint32_t g = 0; g < n; ++g) {
            done += dn[g];
            if (rcs[g] && (rc == VR_OK || rc == VR_ERR_CANCELLED)) {
                rc = rcs[g];
                g_error = errs[g];
            }
        }
        // a cancelled group keeps whatever whole batches each device finished; the pieces are no longer one
        // contiguous range, which only matters to a caller that resumes after a cancel (the reference does not)
    }
    {
        std::lock_guard<std::mutex> lock(r->stats_mutex);
        r->samples_done += done;
    }
    return rc;
} VR_CATCH

int32_t vr_render_cancel(vr_render* r) try {
    if (!r) return fail(VR_ERR_INVALID, "null render");
    if (r->running.load()) {
        r->cancel = 1;
        for (vr_render* sh : r->shards) sh->cancel = 1;
    }
    return VR_OK;
} VR_CATCH

int32_t vr_render_stats(vr_render* r, vr_stats* out) try {
    if (!r || !out) return fail(VR_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(r->stats_mutex);
    out->samples_done = r->samples_done;
    out->total_samples = r->settings.total_samples;
    out->camera_samples = (uint64_t)r->n_pixels * r->samples_done;
    out->ray_segments = r->segments_host;
    out->seconds = r->seconds;
    out->device_ms = r->device_ms;
    out->trace_ms = r->trace_ms;
    out->trace_launches = r->trace_launches;
    out->kernel_launches = r->kernel_launches;
    for (vr_render* sh : r->shards) {  // device group: work adds up, device time is the slowest shard's
        std::lock_guard<std::mutex> shard_lock(sh->stats_mutex);
        out->ray_segments += sh->segments_host;
        out->device_ms = std::max(out->device_ms, sh->device_ms);
        out->trace_ms += sh->trace_ms;
        out->trace_launches += sh->trace_launches;
        out->kernel_launches += sh->kernel_launches;
    }
    return VR_OK;
} VR_CATCH

// Device group: own + shard 1 + shard 2 + ... (that order) per pixel on the first device, the shards read over peer
// memory (NVLink / NVSwitch) inside the kernel, or from staged copies where a device cannot be mapped.
// tonemap < 0: the plain sum goes to `out`; otherwise the resolved pixel (fused reduce + PostProcessingPass).
static int32_t group_reduce(vr_render* r, float4* out, float scale, float gamma, float exposure, int32_t tonemap) {
    vr_context* ctx = r->scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    PeerList peers;
    peers.n = (uint32_t)r->shards.size();
    for (uint32_t k = 0; k < (uint32_t)MAX_PEERS; ++k) peers.ptr[k] = nullptr;
    for (size_t i = 0; i < r->shards.size(); ++i) {
        vr_render* sh = r->shards[i];
        if (ctx->peer_ok[i]) {
            peers.ptr[i] = sh->accum;
        } else {
            VR_CUDA(cudaMemcpyPeerAsync(r->staged[i], ctx->device, sh->accum, sh->scene->ctx->device, 16ull * r->n_pixels, ctx->stream));
            peers.ptr[i] = r->staged[i];
        }
    }
    launch_reduce_resolve_peers(r->accum, peers, out, r->n_pixels, scale, tonemap < 0 ? 1.0f : std::pow(2.0f, exposure),
                                tonemap < 0 ? 1.0f : 1.0f / gamma, tonemap, ctx->stream);
    r->kernel_launches += 1;
    return VR_OK;
}

int32_t vr_render_read_accum(vr_render* r, float* rgba) try {
    if (!r || !rgba) return fail(VR_ERR_INVALID, "null argument");
    cudaStream_t st = r->scene->ctx->stream;
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    const float4* src = r->accum;
    if (!r->shards.empty()) {
        const int32_t rc = group_reduce(r, r->reduced, 1.0f, 1.0f, 0.0f, -1);
        if (rc) return rc;
        src = r->reduced;
    }
    VR_CUDA(cudaMemcpyAsync(rgba, src, 16ull * r->n_pixels, cudaMemcpyDeviceToHost, st));
    VR_CUDA(cudaStreamSynchronize(st));
    VR_CUDA(cudaGetLastError());
    return VR_OK;
} VR_CATCH

int32_t vr_render_accum_device_ptr(vr_render* r, void** device_ptr) try {
    if (!r || !device_ptr) return fail(VR_ERR_INVALID, "null argument");
    *device_ptr = r->accum;
    return VR_OK;
} VR_CATCH

int32_t vr_render_resolve(vr_render* r, float scale, float gamma, float exposure, int32_t tonemap, float* rgba_out) try {
    if (!r || !rgba_out) return fail(VR_ERR_INVALID, "null argument");
    if (tonemap < 0 || tonemap > 4) return fail(VR_ERR_INVALID, "unknown tonemap");
    cudaStream_t st = r->scene->ctx->stream;
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    if (!r->shards.empty()) {
        const int32_t rc = group_reduce(r, r->resolved, scale, gamma, exposure, tonemap);
        if (rc) return rc;
    } else {
        launch_resolve(r->accum, r->resolved, r->n_pixels, scale, std::pow(2.0f, exposure), 1.0f / gamma, tonemap, st);
        r->kernel_launches += 1;
    }
    VR_CUDA(cudaMemcpyAsync(rgba_out, r->resolved, 16ull * r->n_pixels, cudaMemcpyDeviceToHost, st));
    VR_CUDA(cudaStreamSynchronize(st));
    VR_CUDA(cudaGetLastError());
    return VR_OK;
} VR_CATCH

// ---- multi-GPU: peer-memory reduce ----------------------------------------------------------------

int32_t vr_render_export_accum(vr_render* r, uint8_t handle[VR_IPC_HANDLE_BYTES]) try {
    if (!r || !handle) return fail(VR_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == VR_IPC_HANDLE_BYTES, "IPC handle size");
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    cudaIpcMemHandle_t h;
    VR_CUDA(cudaIpcGetMemHandle(&h, r->accum));
    std::memcpy(handle, &h, VR_IPC_HANDLE_BYTES);
    return VR_OK;
} VR_CATCH

static int32_t peer_reduce(vr_render* r, void* const* peer_accum, uint32_t n_peers, float scale, float gamma,
                           float exposure, int32_t tonemap, float* rgba_out) {
    if (n_peers > (uint32_t)MAX_PEERS) return fail(VR_ERR_INVALID, "too many peers (max 15)");
    if (n_peers && !peer_accum) return fail(VR_ERR_INVALID, "null peer list");
    cudaStream_t st = r->scene->ctx->stream;
    PeerList peers;
    peers.n = n_peers;
    for (uint32_t k = 0; k < n_peers; ++k) {
        if (!peer_accum[k]) return fail(VR_ERR_INVALID, "null peer pointer");
        peers.ptr[k] = (const float4*)peer_accum[k];
    }
    for (uint32_t k = n_peers; k < (uint32_t)MAX_PEERS; ++k) peers.ptr[k] = nullptr;
    if (tonemap < 0) {
        launch_reduce_resolve_peers(r->accum, peers, r->accum, r->n_pixels, 1.0f, 1.0f, 1.0f, -1, st);
        r->kernel_launches += 1;
        VR_CUDA(cudaStreamSynchronize(st));
    } else {
        launch_reduce_resolve_peers(r->accum, peers, r->resolved, r->n_pixels, scale, std::pow(2.0f, exposure), 1.0f / gamma,
                                    tonemap, st);
        r->kernel_launches += 1;
        VR_CUDA(cudaMemcpyAsync(rgba_out, r->resolved, 16ull * r->n_pixels, cudaMemcpyDeviceToHost, st));
        VR_CUDA(cudaStreamSynchronize(st));
    }
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

static int32_t open_peers(vr_render* r, const uint8_t* peer_handles, uint32_t n_peers, std::vector<void*>& out) {
    if (n_peers && !peer_handles) return fail(VR_ERR_INVALID, "null peer handles");
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    out.clear();
    for (uint32_t k = 0; k < n_peers; ++k) {
        const std::string key((const char*)peer_handles + (size_t)k * VR_IPC_HANDLE_BYTES, VR_IPC_HANDLE_BYTES);
        void* p = nullptr;
        for (auto& e : r->ipc_peers)
            if (e.first == key) p = e.second;
        if (!p) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, key.data(), VR_IPC_HANDLE_BYTES);
            VR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            r->ipc_peers.emplace_back(key, p);
        }
        out.push_back(p);
    }
    return VR_OK;
}

int32_t vr_render_reduce_peers(vr_render* r, const uint8_t* peer_handles, uint32_t n_peers) try {
    if (!r) return fail(VR_ERR_INVALID, "null render");
    std::vector<void*> ptrs;
    const int32_t rc = open_peers(r, peer_handles, n_peers, ptrs);
    if (rc) return rc;
    return peer_reduce(r, ptrs.data(), n_peers, 1.0f, 1.0f, 0.0f, -1, nullptr);
} VR_CATCH

int32_t vr_render_resolve_peers(vr_render* r, const uint8_t* peer_handles, uint32_t n_peers, float scale, float gamma,
                                float exposure, int32_t tonemap, float* rgba_out) try {
    if (!r || !rgba_out) return fail(VR_ERR_INVALID, "null argument");
    if (tonemap < 0 || tonemap > 4) return fail(VR_ERR_INVALID, "unknown tonemap");
    std::vector<void*> ptrs;
    const int32_t rc = open_peers(r, peer_handles, n_peers, ptrs);
    if (rc) return rc;
    return peer_reduce(r, ptrs.data(), n_peers, scale, gamma, exposure, tonemap, rgba_out);
} VR_CATCH

int32_t vr_render_reduce_peer_ptrs(vr_render* r, void* const* peer_accum, uint32_t n_peers) try {
    if (!r) return fail(VR_ERR_INVALID, "null render");
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    return peer_reduce(r, peer_accum, n_peers, 1.0f, 1.0f, 0.0f, -1, nullptr);
} VR_CATCH

int32_t vr_render_resolve_peer_ptrs(vr_render* r, void* const* peer_accum, uint32_t n_peers, float scale, float gamma,
                                    float exposure, int32_t tonemap, float* rgba_out) try {
    if (!r || !rgba_out) return fail(VR_ERR_INVALID, "null argument");
    if (tonemap < 0 || tonemap > 4) return fail(VR_ERR_INVALID, "unknown tonemap");
    VR_CUDA(cudaSetDevice(r->scene->ctx->device));
    return peer_reduce(r, peer_accum, n_peers, scale, gamma, exposure, tonemap, rgba_out);
} VR_CATCH

// ---- gates ----------------------------------------------------------------------------------------

int32_t vr_debug_trace_primary(vr_render* r, uint32_t sample, uint32_t* surface, uint32_t* prim, float* t) try {
    if (!r || !surface || !prim || !t) return fail(VR_ERR_INVALID, "null argument");
    vr_context* ctx = r->scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    const FrameParams fp = frame_params(r);
    const PathSource src = make_path_source(nullptr, nullptr, r->width, r->height, sample);
    VR_CUDA(cudaMemsetAsync(r->wf.counts, 0, sizeof(uint32_t) * 2 * (fp.max_bounces + 2), ctx->stream));
    launch_raygen(r->scene->dev, r->wf, src, fp, r->n_pixels, false, ctx->dims, ctx->stream);
    launch_trace(r->scene->dev, r->wf, 0, r->n_pixels, ctx->dims, ctx->stream);
    launch_primary_ids(r->scene->dev, r->wf, r->width, r->height, r->dbg_surface, r->dbg_prim, r->dbg_t, ctx->stream);
    VR_CUDA(cudaMemcpyAsync(surface, r->dbg_surface, 4ull * r->n_pixels, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaMemcpyAsync(prim, r->dbg_prim, 4ull * r->n_pixels, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaMemcpyAsync(t, r->dbg_t, 4ull * r->n_pixels, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    VR_CUDA(cudaGetLastError());
    return VR_OK;
} VR_CATCH

int32_t vr_debug_trace_rays(vr_scene* scene, uint64_t n, const float* origins, const float* directions,
                            uint32_t* surface, uint32_t* prim, float* t) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!scene->committed) return fail(VR_ERR_INVALID, "scene not committed");
    if (n == 0) return VR_OK;
    if (!origins || !directions || !surface || !prim || !t) return fail(VR_ERR_INVALID, "null argument");
    vr_context* ctx = scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    DeviceBuffers tmp;
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
    uint32_t *d_s = nullptr, *d_p = nullptr;
    cudaError_t e = tmp.alloc((void**)&d_o, 12 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_d, 12 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_t, 4 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_s, 4 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_p, 4 * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_o, origins, 12 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, directions, 12 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        launch_trace_rays(scene->dev, d_o, d_d, n, d_s, d_p, d_t, ctx->stream);
        e = cudaMemcpyAsync(surface, d_s, 4 * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(prim, d_p, 4 * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t, d_t, 4 * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    tmp.release();
    if (e != cudaSuccess) return fail(VR_ERR_CUDA, std::string("vr_debug_trace_rays: ") + cudaGetErrorString(e));
    return VR_OK;
} VR_CATCH

int32_t vr_debug_sample_radiance(vr_render* r, uint64_t n, const uint32_t* pixel, const uint32_t* sample, float* out) try {
    if (!r || !pixel || !sample || !out) return fail(VR_ERR_INVALID, "null argument");
    if (!r->scene->committed) return fail(VR_ERR_INVALID, "scene was edited after commit");
    vr_context* ctx = r->scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    if (r->settings.integrator == 1) {
        const int32_t rc = ensure_env_tables(r->scene);
        if (rc) return rc;
    }
    for (uint64_t i = 0; i < n; ++i)
        if (pixel[i] >= r->n_pixels) return fail(VR_ERR_INVALID, "pixel index out of range");
    DeviceBuffers tmp;
    const uint64_t chunk_max = r->wf.capacity;
    uint32_t *d_px = nullptr, *d_sm = nullptr;
    cudaError_t e = tmp.alloc((void**)&d_px, 4 * chunk_max);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_sm, 4 * chunk_max);
    std::vector<float4> host(chunk_max < n ? chunk_max : n);
    for (uint64_t base = 0; base < n && e == cudaSuccess; base += chunk_max) {
        const uint32_t m = (uint32_t)std::min<uint64_t>(chunk_max, n - base);
        e = cudaMemcpyAsync(d_px, pixel + base, 4ull * m, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_sm, sample + base, 4ull * m, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) break;
        const PathSource src = make_path_source(d_px, d_sm, r->width, r->height, 0);
        size_t cursor = 0;
        run_wavefront(r, r->wf, ctx->stream, src, m, false, &cursor);
        e = cudaMemcpyAsync(host.data(), r->wf.radiance, 16ull * m, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        for (uint32_t i = 0; i < m && e == cudaSuccess; ++i) {
            out[3 * (base + i)] = host[i].x;
            out[3 * (base + i) + 1] = host[i].y;
            out[3 * (base + i) + 2] = host[i].z;
        }
    }
    tmp.release();
    if (e != cudaSuccess) return fail(VR_ERR_CUDA, std::string("vr_debug_sample_radiance: ") + cudaGetErrorString(e));
    return VR_OK;
} VR_CATCH

int32_t vr_debug_reference_leaf_order(const float* boxes6, uint64_t n, uint32_t* order) try {
    if ((!boxes6 || !order) && n) return fail(VR_ERR_INVALID, "null argument");
    if (n >= (1ull << 32)) return fail(VR_ERR_INVALID, "too many items");
    RawVector<uint32_t> o;
    reference_leaf_order(boxes6, (size_t)n, o);
    if (n) std::memcpy(order, o.data(), (size_t)n * sizeof(uint32_t));
    return VR_OK;
} VR_CATCH

int32_t vr_debug_flatten_mesh_digest(const float* positions, const float* uvs, const float* normals, uint32_t n_vertices,
                                     const uint32_t* indices, uint32_t n_indices, uint64_t* digest, uint32_t* n_nodes,
                                     uint32_t* bvh_depth, double* flatten_ms) try {
    if (!positions || !indices || !digest) return fail(VR_ERR_INVALID, "null argument");
    HostScene sc;
    HostMesh m;
    m.n_vertices = n_vertices;
    m.pos.assign(positions, positions + 3ull * n_vertices);
    if (uvs) m.uv.assign(uvs, uvs + 2ull * n_vertices);
    else m.uv.assign(2ull * n_vertices, 0.0f);
    if (normals) m.nrm.assign(normals, normals + 3ull * n_vertices);
    else m.nrm.assign(3ull * n_vertices, 0.0f);
    const uint32_t n_idx = n_indices - n_indices % 3;
    for (uint32_t i = 0; i < n_idx; ++i)
        if (indices[i] >= n_vertices) return fail(VR_ERR_INVALID, "mesh index out of range");
    m.idx.assign(indices, indices + n_idx);
    sc.meshes.push_back(std::move(m));
    HostSurface sf;
    sf.kind = 0;
    sf.mesh = 0;
    sc.surfaces.push_back(sf);
    MaterialRec mat{};
    mat.albedo_tex = -1;
    mat.normal_tex = -1;
    sc.materials.push_back(mat);
    sc.objects.push_back(HostObject{0, 0});
    const float eye[3] = {0.0f, 0.0f, -1.0f}, dir[3] = {0.0f, 0.0f, 1.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    std::memcpy(sc.camera.eye, eye, 12);
    std::memcpy(sc.camera.direction, dir, 12);
    std::memcpy(sc.camera.up, up, 12);
    sc.camera.fov = 0.6f;
    sc.camera.has_dof = 0;
    FlatScene flat;
    std::string err;
    const auto t0 = std::chrono::steady_clock::now();
    if (!flatten_scene(sc, flat, err)) return fail(VR_ERR_INVALID, err);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    auto fnv = [](const void* ptr, size_t n, uint64_t d) {  // FNV-1a
        const unsigned char* b = (const unsigned char*)ptr;
        for (size_t i = 0; i < n; ++i) d = (d ^ b[i]) * 1099511628211ull;
        return d;
    };
    uint64_t d = 1469598103934665603ull;
    d = fnv(flat.nodes.data(), flat.nodes.size() * sizeof(Quad), d);
    d = fnv(flat.tri_isect.data(), flat.tri_isect.size() * sizeof(Quad), d);
    d = fnv(flat.tri_shade.data(), flat.tri_shade.size() * sizeof(Quad), d);
    *digest = d;
    if (n_nodes) *n_nodes = (uint32_t)(flat.nodes.size() / DEVICE_NODE_QUADS);
    if (bvh_depth) *bvh_depth = flat.bvh_depth;
    if (flatten_ms) *flatten_ms = ms;
    return VR_OK;
} VR_CATCH

int32_t vr_debug_tie_ranks(vr_scene* scene, uint32_t surface, uint32_t* out, uint32_t n) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!scene->committed) return fail(VR_ERR_INVALID, "scene not committed");
    if (surface >= scene->flat.mesh_tie_rank.size()) return fail(VR_ERR_INVALID, "unknown surface");
    const RawVector<uint32_t>& rank = scene->flat.mesh_tie_rank[surface];
    if (n != rank.size()) return fail(VR_ERR_INVALID, "n must equal the surface's triangle count");
    if (n && !out) return fail(VR_ERR_INVALID, "null argument");
    for (uint32_t i = 0; i < n; ++i) out[i] = scene->flat.surface_rank_base[surface] + rank[i];
    return VR_OK;
} VR_CATCH

int32_t vr_debug_texture_sample(vr_scene* scene, uint32_t texture, uint64_t n, const float* uv, float* rgb) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!scene->committed) return fail(VR_ERR_INVALID, "scene not committed");
    if (texture >= scene->dev_textures.size()) return fail(VR_ERR_INVALID, "unknown texture");
    if (n == 0) return VR_OK;
    if (!uv || !rgb) return fail(VR_ERR_INVALID, "null argument");
    vr_context* ctx = scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    DeviceBuffers tmp;
    float *d_uv = nullptr, *d_rgb = nullptr;
    cudaError_t e = tmp.alloc((void**)&d_uv, 8 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_rgb, 12 * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_uv, uv, 8 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        launch_texture_sample(scene->dev_textures[texture], n, d_uv, d_rgb, ctx->stream);
        e = cudaMemcpyAsync(rgb, d_rgb, 12 * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    tmp.release();
    if (e != cudaSuccess) return fail(VR_ERR_CUDA, std::string("vr_debug_texture_sample: ") + cudaGetErrorString(e));
    return VR_OK;
} VR_CATCH

int32_t vr_debug_environment_sample(vr_scene* scene, uint64_t n, const float* directions, float* rgb) try {
    if (check_scene(scene)) return VR_ERR_INVALID;
    if (!scene->committed) return fail(VR_ERR_INVALID, "scene not committed");
    if (n == 0) return VR_OK;
    if (!directions || !rgb) return fail(VR_ERR_INVALID, "null argument");
    vr_context* ctx = scene->ctx;
    VR_CUDA(cudaSetDevice(ctx->device));
    DeviceBuffers tmp;
    float *d_d = nullptr, *d_rgb = nullptr;
    cudaError_t e = tmp.alloc((void**)&d_d, 12 * n);
    if (e == cudaSuccess) e = tmp.alloc((void**)&d_rgb, 12 * n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, directions, 12 * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        launch_environment_sample(scene->dev, n, d_d, d_rgb, ctx->stream);
        e = cudaMemcpyAsync(rgb, d_rgb, 12 * n, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    tmp.release();
    if (e != cudaSuccess) return fail(VR_ERR_CUDA, std::string("vr_debug_environment_sample: ") + cudaGetErrorString(e));
    return VR_OK;
} VR_CATCH

static int32_t debug_draws(vr_context* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, void* out,
                           int floats_per_item) {
    if (!ctx || !out) return fail(VR_ERR_INVALID, "null argument");
    if (n == 0) return VR_OK;
    VR_CUDA(cudaSetDevice(ctx->device));
    void* d = nullptr;
    const size_t bytes = 4ull * n * floats_per_item;
    VR_CUDA(cudaMalloc(&d, bytes));
    if (floats_per_item == 1) launch_rng_draws(seed, pixel, sample, n, (uint32_t*)d, ctx->stream);
    else launch_unit_sphere(seed, pixel, sample, n, (float*)d, ctx->stream);
    cudaError_t e = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d);
    if (e != cudaSuccess) return fail(VR_ERR_CUDA, std::string("debug draws: ") + cudaGetErrorString(e));
    return VR_OK;
}

int32_t vr_debug_rng_draws(vr_context* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, uint32_t* out) try {
    return debug_draws(ctx, seed, pixel, sample, n, out, 1);
} VR_CATCH
int32_t vr_debug_unit_sphere(vr_context* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out) try {
    return debug_draws(ctx, seed, pixel, sample, n, out, 3);
} VR_CATCH

}  // extern "C"
