// voidray.hpp — header-only C++ mirror of the reference renderer's host API over the C ABI (voidray_cuda.h).
//
// The reference is Rust; this is the compiled-language host side a port of voidray_app would be written against.
// Names, argument meaning and error behaviour follow the reference (paths relative to its checkout):
//   Scene / handles            voidray_renderer/src/core/scene.rs:36-161
//   Camera, Camera::look_at    voidray_renderer/src/core/camera.rs:7-36
//   Materials, MicrofacetBSDF  voidray_common/src/simple.rs:17-58, microfacet.rs:29-101
//   Surfaces                   voidray_common/src/surfaces.rs:9-29
//   Environments               voidray_common/src/environments.rs:9-17
//   Settings                   voidray_renderer/src/core/settings.rs
//   build_acceleration         core/scene.rs:163-179          -> SceneAcceleration (device-resident)
//   CpuRenderTarget            render/target.rs:80-299        -> RenderTarget (accumulation buffer in HBM)
//   iterative_render           render/iterative.rs:11-55
//   PostProcessingPass         render/post_process.rs:19-86
//   Renderer / RenderAction    render/renderer.rs:35-258      (RenderThread::one_shot on a std::thread)
// Where the reference panics (unwrap on I/O errors, invalid actions) this mirror throws voidray::Error.
#pragma once

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <variant>
#include <vector>

#include "voidray_cuda.h"

namespace voidray {

struct Error : std::runtime_error {
    int32_t status;
    Error(int32_t s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(int32_t status) {
    if (status != VR_OK) throw Error(status, std::string("voidray_cuda status ") + std::to_string(status) + ": " + vr_last_error());
}

typedef float Float;  // util/vector.rs:11
struct Vec3 {
    Float x, y, z;
};
struct Color {
    Float r, g, b;
};
// util/color.rs:18-23
inline Color hex_color(uint32_t x) {
    return Color{(Float)((x >> 16) & 0xff) / 255.0f, (Float)((x >> 8) & 0xff) / 255.0f, (Float)(x & 0xff) / 255.0f};
}

// Opaque index newtypes, core/scene.rs:46-59
struct MaterialHandle { uint32_t v; };
struct ObjectHandle { uint32_t v; };
struct SurfaceHandle { uint32_t v; };
struct TextureHandle { uint32_t v; };

enum class SampleType { Nearest = 0, Bilinear = 1 };                                   // core/texture.rs:23-26
enum class RenderMode { Full = 0, Normal = 1 };                                        // core/settings.rs:9-13
enum class Tonemap { None = 0, ACES = 1, Reinhard = 2, Filmic = 3, Uncharted2 = 4 };   // core/settings.rs:36-55

struct RenderSettings {  // core/settings.rs:15-33
    uint32_t total_samples = 100;
    float update_frequency = 0.1f;
    RenderMode render_mode = RenderMode::Full;
    Float firefly_clamp = 3.0f;
    uint32_t max_bounces = 10;
    // not in the reference (see vr_render_settings)
    uint64_t seed = 0x5EED0001ull;
    int32_t pixel_mapping = 0;
    int32_t integrator = 0;
    uint32_t sample_offset = 0;
};
struct ColorManagementSettings {  // core/settings.rs:57-74
    Tonemap tonemap = Tonemap::None;
    float gamma = 2.2f;
    float exposure = 0.0f;
    bool transparent = true;
};
struct Settings {
    RenderSettings render;
    ColorManagementSettings color_management;
};

struct Camera {  // core/camera.rs:7-22
    Vec3 eye{1, 0, 10}, direction{0, 0, -1}, up{0, 1, 0};
    Float fov = 0.5235988f;
    bool has_dof = false;
    Float aperture = 0;
    Vec3 focal_point{0, 0, 0};
    // Camera::look_at, camera.rs:26-36 — direction and up are computed here, once, by the library's own routine
    // (the reference's f32 operation order), so later edits of `eye` leave them alone exactly like in the reference
    static Camera look_at(Vec3 eye, Vec3 center, Vec3 up, Float fov) {
        Camera c;
        const float e[3] = {eye.x, eye.y, eye.z}, ce[3] = {center.x, center.y, center.z}, u[3] = {up.x, up.y, up.z};
        float d[3], u2[3];
        check(vr_camera_look_at(e, ce, u, d, u2));
        c.eye = eye;
        c.direction = Vec3{d[0], d[1], d[2]};
        c.up = Vec3{u2[0], u2[1], u2[2]};
        c.fov = fov;
        return c;
    }
};

// ---- materials: the closed set of `dyn Material` implementations ------------------------------------
struct Materials {  // voidray_common/src/simple.rs:17-58
    static vr_material_desc make(int32_t kind, Color c, float param, int32_t albedo_tex = -1, int32_t normal_tex = -1) {
        vr_material_desc d{};
        d.kind = kind;
        d.color[0] = c.r; d.color[1] = c.g; d.color[2] = c.b;
        d.param = param;
        d.albedo_tex = albedo_tex;
        d.normal_tex = normal_tex;
        return d;
    }
    static vr_material_desc lambertian(Color albedo) { return make(VR_MAT_LAMBERTIAN, albedo, 0); }
    static vr_material_desc lambertian_bsdf(Color albedo) { return make(VR_MAT_LAMBERTIAN_BSDF, albedo, 0); }
    static vr_material_desc lambertian_texture_no_normal(TextureHandle albedo) { return make(VR_MAT_LAMBERTIAN, Color{0, 0, 0}, 0, (int32_t)albedo.v); }
    static vr_material_desc lambertian_texture(TextureHandle albedo, TextureHandle normal) {
        return make(VR_MAT_LAMBERTIAN, Color{0, 0, 0}, 0, (int32_t)albedo.v, (int32_t)normal.v);
    }
    static vr_material_desc metal(Color albedo, Float fuzz) { return make(VR_MAT_METAL, albedo, fuzz); }
    static vr_material_desc dielectric(Float ir) { return make(VR_MAT_DIELECTRIC, Color{0, 0, 0}, ir); }
    static vr_material_desc emissive(Float strength) { return make(VR_MAT_EMISSION, Color{1, 1, 1}, strength); }
    static vr_material_desc colored_emissive(Color color, Float strength) { return make(VR_MAT_EMISSION, color, strength); }
};
struct MicrofacetBSDF {  // voidray_common/src/microfacet.rs:29-101
    static vr_material_desc make(Color c, Float index, Float roughness, Float metallic, Float emittance, bool transparent) {
        vr_material_desc d = Materials::make(VR_MAT_MICROFACET, c, 0);
        d.index = index; d.roughness = roughness; d.metallic = metallic; d.emittance = emittance;
        d.transparent = transparent ? 1 : 0;
        return d;
    }
    static vr_material_desc diffuse(Color c) { return make(c, 1.5f, 1.0f, 0, 0, false); }
    static vr_material_desc specular(Color c, Float roughness) { return make(c, 1.5f, roughness, 0, 0, false); }
    static vr_material_desc clear(Float index, Float roughness) { return make(hex_color(0xFFFFFF), index, roughness, 0, 0, true); }
    static vr_material_desc transparent(Color c, Float index, Float roughness) { return make(c, index, roughness, 0, 0, true); }
    static vr_material_desc metallic(Color c, Float roughness) { return make(c, 1.5f, roughness, 1.0f, 0, false); }
    static vr_material_desc light(Color c, Float emittance) { return make(c, 1.0f, 1.0f, 0, emittance, false); }
};

// ---- surfaces -------------------------------------------------------------------------------------
struct Mesh {  // core/mesh.rs:35-41 before acceleration
    std::vector<float> positions, uvs, normals;  // 3n, 2n, 3n
    std::vector<uint32_t> indices;
    // Vertex::position (mesh.rs:26-32): uv and normal are zero
    static Mesh from_positions(const std::vector<Vec3>& p, std::vector<uint32_t> idx) {
        Mesh m;
        for (const Vec3& v : p) { m.positions.push_back(v.x); m.positions.push_back(v.y); m.positions.push_back(v.z); }
        m.uvs.assign(2 * p.size(), 0.0f);
        m.normals.assign(3 * p.size(), 0.0f);
        m.indices = std::move(idx);
        return m;
    }
};
struct Sphere { Vec3 center; Float radius; };
struct GroundPlane { Float height; };
struct ObjFile { std::string path; };
struct Surfaces {  // voidray_common/src/surfaces.rs:9-29
    static Sphere sphere(Vec3 center, Float radius) { return Sphere{center, radius}; }
    static GroundPlane ground_plane(Float height) { return GroundPlane{height}; }
    static Mesh quad(Vec3 q1, Vec3 q2, Vec3 q3, Vec3 q4) { return Mesh::from_positions({q1, q2, q3, q4}, {0, 1, 2, 2, 0, 3}); }
};

// ---- environments ---------------------------------------------------------------------------------
// image::open(path).unwrap().to_rgb32f() (core/texture.rs:37, environments.rs:43) through the library's decoders;
// throws where the reference panics.
inline std::vector<float> load_image_rgb32f(const std::string& path, uint32_t* width, uint32_t* height) {
    float* rgb = nullptr;
    check(vr_image_load_rgb32f(path.c_str(), width, height, &rgb));
    std::vector<float> out(rgb, rgb + (size_t)3 * *width * *height);
    vr_image_free(rgb);
    return out;
}

struct UniformEnvironment { Color color; };
struct HDRIEnvironment { std::vector<float> rgb; uint32_t width, height; };
struct NoEnvironment {};
typedef std::variant<NoEnvironment, UniformEnvironment, HDRIEnvironment> Environment;
struct Environments {  // voidray_common/src/environments.rs:9-17 (the image arrives decoded: to_rgb32f)
    static Environment uniform(Color background) { return UniformEnvironment{background}; }
    static Environment hdri(std::vector<float> rgb, uint32_t width, uint32_t height) { return HDRIEnvironment{std::move(rgb), width, height}; }
    static Environment hdri(const std::string& path) {  // environments.rs:13-16,42-55
        uint32_t w = 0, h = 0;
        std::vector<float> rgb = load_image_rgb32f(path, &w, &h);
        return HDRIEnvironment{std::move(rgb), w, h};
    }
};

struct ImageTexture {
    std::vector<float> rgb;
    uint32_t width, height;
    SampleType sample_type;
};

// ---- context ----------------------------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0, void* cuda_stream = nullptr) { check(vr_context_create(device, cuda_stream, &ctx_)); }
    // a device group in this process: commit replicates, accumulate shards the samples, read / resolve sum over peer memory
    explicit Context(const std::vector<int32_t>& devices) {
        check(vr_context_create_multi(devices.data(), (uint32_t)devices.size(), &ctx_));
    }
    ~Context() { vr_context_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    vr_context* raw() const { return ctx_; }
private:
    vr_context* ctx_ = nullptr;
};

class SceneAcceleration;

// ---- Scene: pure host data, like the reference's (core/scene.rs:36-44) --------------------------------
class Scene {
public:
    Camera camera = Camera::look_at(Vec3{1, 0, 10}, Vec3{0, 0, 0}, Vec3{0, 1, 0}, 3.14159265358979323846f / 6.0f);  // scene.rs:95-111
    Environment environment = NoEnvironment{};

    static Scene empty() { return Scene(); }
    MaterialHandle add_material(const vr_material_desc& m) { materials_.push_back(m); return MaterialHandle{(uint32_t)materials_.size() - 1}; }
    SurfaceHandle add_analytic_surface(Sphere s) { surfaces_.emplace_back(s); return SurfaceHandle{(uint32_t)surfaces_.size() - 1}; }
    SurfaceHandle add_analytic_surface(GroundPlane g) { surfaces_.emplace_back(g); return SurfaceHandle{(uint32_t)surfaces_.size() - 1}; }
    SurfaceHandle add_mesh(Mesh m) { surfaces_.emplace_back(std::move(m)); return SurfaceHandle{(uint32_t)surfaces_.size() - 1}; }
    SurfaceHandle add_mesh_from_file(const std::string& path) { surfaces_.emplace_back(ObjFile{path}); return SurfaceHandle{(uint32_t)surfaces_.size() - 1}; }
    ObjectHandle add_object(MaterialHandle material, SurfaceHandle surface) {
        objects_.emplace_back(material.v, surface.v);
        return ObjectHandle{(uint32_t)objects_.size() - 1};
    }
    TextureHandle add_image_texture(std::vector<float> rgb, uint32_t width, uint32_t height, SampleType sample_type) {
        textures_.push_back(ImageTexture{std::move(rgb), width, height, sample_type});
        return TextureHandle{(uint32_t)textures_.size() - 1};
    }
    TextureHandle add_image_texture(const std::string& path, SampleType sample_type) {  // scene.rs:154-160
        uint32_t w = 0, h = 0;
        std::vector<float> rgb = load_image_rgb32f(path, &w, &h);
        return add_image_texture(std::move(rgb), w, h, sample_type);
    }
    // Accelerable::build_acceleration, core/scene.rs:163-179
    std::shared_ptr<SceneAcceleration> build_acceleration(const Context& ctx) const;

private:
    friend class SceneAcceleration;
    typedef std::variant<Mesh, ObjFile, Sphere, GroundPlane> Surface;
    std::vector<vr_material_desc> materials_;
    std::vector<Surface> surfaces_;
    std::vector<std::pair<uint32_t, uint32_t>> objects_;  // (material, surface)
    std::vector<ImageTexture> textures_;
};

// ---- SceneAcceleration: the committed, device-resident scene ------------------------------------------
class SceneAcceleration {
public:
    SceneAcceleration(const Context& ctx, const Scene& s) {
        check(vr_scene_create(ctx.raw(), &scene_));
        try {
            uint32_t out;
            for (const ImageTexture& t : s.textures_)
                check(vr_scene_add_texture_rgb32f(scene_, t.rgb.data(), t.width, t.height, (int32_t)t.sample_type, &out));
            for (const Scene::Surface& sf : s.surfaces_) {
                if (const Mesh* m = std::get_if<Mesh>(&sf))
                    check(vr_scene_add_mesh(scene_, m->positions.data(), m->uvs.data(), m->normals.data(), (uint32_t)(m->positions.size() / 3),
                                            m->indices.data(), (uint32_t)m->indices.size(), &out));
                else if (const ObjFile* o = std::get_if<ObjFile>(&sf))
                    check(vr_scene_add_mesh_from_obj_file(scene_, o->path.c_str(), &out, nullptr, nullptr));
                else if (const Sphere* sp = std::get_if<Sphere>(&sf)) {
                    const float c[3] = {sp->center.x, sp->center.y, sp->center.z};
                    check(vr_scene_add_sphere(scene_, c, sp->radius, &out));
                } else
                    check(vr_scene_add_ground_plane(scene_, std::get<GroundPlane>(sf).height, &out));
            }
            for (const vr_material_desc& m : s.materials_) check(vr_scene_add_material(scene_, &m, &out));
            for (const auto& o : s.objects_) check(vr_scene_add_object(scene_, o.first, o.second, &out));
            const Camera& c = s.camera;
            const float eye[3] = {c.eye.x, c.eye.y, c.eye.z}, up[3] = {c.up.x, c.up.y, c.up.z};
            const float dir[3] = {c.direction.x, c.direction.y, c.direction.z}, fp[3] = {c.focal_point.x, c.focal_point.y, c.focal_point.z};
            check(vr_scene_set_camera(scene_, eye, dir, up, c.fov, c.has_dof ? 1 : 0, c.aperture, fp));
            if (const UniformEnvironment* u = std::get_if<UniformEnvironment>(&s.environment)) {
                const float rgb[3] = {u->color.r, u->color.g, u->color.b};
                check(vr_scene_set_environment_uniform(scene_, rgb));
            } else if (const HDRIEnvironment* h = std::get_if<HDRIEnvironment>(&s.environment)) {
                check(vr_scene_set_environment_hdri_rgb32f(scene_, h->rgb.data(), h->width, h->height));
            }
            check(vr_scene_commit(scene_));
        } catch (...) {
            vr_scene_destroy(scene_);
            throw;
        }
    }
    ~SceneAcceleration() { vr_scene_destroy(scene_); }
    SceneAcceleration(const SceneAcceleration&) = delete;
    SceneAcceleration& operator=(const SceneAcceleration&) = delete;
    vr_scene* raw() const { return scene_; }
    vr_scene_info info() const {
        vr_scene_info i;
        check(vr_scene_get_info(scene_, &i));
        return i;
    }
private:
    vr_scene* scene_ = nullptr;
};

inline std::shared_ptr<SceneAcceleration> Scene::build_acceleration(const Context& ctx) const {
    return std::make_shared<SceneAcceleration>(ctx, *this);
}

// ---- RenderTarget: the accumulation buffer (CpuRenderTarget semantics), resident in HBM -----------------
class RenderTarget {
public:
    RenderTarget(std::shared_ptr<SceneAcceleration> scene, uint32_t width, uint32_t height, const RenderSettings& s)
        : scene_(std::move(scene)), width_(width), height_(height) {
        vr_render_settings rs{};
        rs.total_samples = s.total_samples;
        rs.max_bounces = s.max_bounces;
        rs.firefly_clamp = s.firefly_clamp;
        rs.render_mode = (int32_t)s.render_mode;
        rs.pixel_mapping = s.pixel_mapping;
        rs.integrator = s.integrator;
        rs.seed = s.seed;
        rs.sample_offset = s.sample_offset;
        rs.max_paths_in_flight = 0;
        check(vr_render_begin(scene_->raw(), width, height, &rs, &render_));
    }
    ~RenderTarget() { vr_render_end(render_); }
    RenderTarget(const RenderTarget&) = delete;
    RenderTarget& operator=(const RenderTarget&) = delete;
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const std::shared_ptr<SceneAcceleration>& scene() const { return scene_; }
    void clear() { check(vr_render_clear(render_)); }                       // target.rs:284-290
    void cancel() { check(vr_render_cancel(render_)); }
    vr_stats stats() const {
        vr_stats st;
        check(vr_render_stats(render_, &st));
        return st;
    }
    std::vector<float> read() const {                                      // W*H*4 partial sums / total_samples
        std::vector<float> out((size_t)width_ * height_ * 4);
        check(vr_render_read_accum(render_, out.data()));
        return out;
    }
    vr_render* raw() const { return render_; }
private:
    std::shared_ptr<SceneAcceleration> scene_;
    uint32_t width_, height_;
    vr_render* render_ = nullptr;
};

// render/iterative.rs:11-55. Returns false if the call was cancelled (vr_render_cancel).
inline bool iterative_render(RenderTarget& target, const SceneAcceleration& scene, const RenderSettings&, uint32_t samples) {
    if (target.scene().get() != &scene) throw Error(VR_ERR_INVALID, "iterative_render: the target was created for another scene");
    const int32_t st = vr_render_accumulate(target.raw(), samples);
    if (st == VR_ERR_CANCELLED) return false;
    check(st);
    return true;
}

struct PostProcessingData {  // render/post_process.rs:19-25
    float scale, gamma, exposure;
    int32_t tonemap;
};
struct PostProcessingPass {  // render/post_process.rs:27-86, over the whole target
    std::vector<float> render(const RenderTarget& src, const PostProcessingData& d) const {
        std::vector<float> out((size_t)src.width() * src.height() * 4);
        check(vr_render_resolve(src.raw(), d.scale, d.gamma, d.exposure, d.tonemap, out.data()));
        return out;
    }
};

enum class RenderAction { Render, Continuous, Rebuild, Cancel };  // render/renderer.rs:165-174

// render/renderer.rs:155-258 with RenderThread::one_shot (:35-125)
class Renderer {
public:
    Renderer(const Context& ctx, std::shared_ptr<Scene> scene, std::shared_ptr<Settings> settings, uint32_t width, uint32_t height)
        : ctx_(ctx), scene_(std::move(scene)), settings_(std::move(settings)), width_(width), height_(height) {}
    ~Renderer() {
        if (thread_.joinable()) {
            cancel_ = true;
            if (target_) vr_render_cancel(target_->raw());
            thread_.join();
        }
    }
    void execute(RenderAction action) {
        if (action == RenderAction::Render) {
            if (currently_rendering()) throw Error(VR_ERR_INVALID, "invalid action Render");  // renderer.rs:203-205 panics
            if (thread_.joinable()) thread_.join();
            cancel_ = false;
            rendering_ = true;
            thread_ = std::thread([this]() { one_shot(); });
        } else if (action == RenderAction::Cancel) {
            if (!thread_.joinable()) throw Error(VR_ERR_INVALID, "invalid action Cancel");    // renderer.rs:229-231
            cancel_ = true;
            std::lock_guard<std::mutex> lock(mutex_);
            if (target_) target_->cancel();
        } else {
            throw Error(VR_ERR_INVALID, "invalid action");  // Continuous / Rebuild are stubs in the reference
        }
    }
    void join() {
        if (thread_.joinable()) thread_.join();
        std::lock_guard<std::mutex> lock(mutex_);
        if (!error_.empty()) throw Error(VR_ERR_CUDA, error_);
    }
    bool currently_rendering() const { return rendering_; }
    std::pair<uint32_t, uint32_t> samples() const {
        std::lock_guard<std::mutex> lock(mutex_);
        return samples_;
    }
    double elapsed_time() const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!started_) return 0.0;
        const auto end = rendering_ ? std::chrono::steady_clock::now() : end_;
        return std::chrono::duration<double>(end - start_).count();
    }
    // RendererStats.remaining (renderer.rs:75-77,95-98): elapsed / done * (total - done); negative = None
    double remaining_time() const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!rendering_ || !started_ || samples_.first == 0) return -1.0;
        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count();
        return elapsed / samples_.first * (samples_.second - samples_.first);
    }
    // voidray_app/src/main.rs:68-90: scale = total / done (0 if not normal), then the tonemap pass
    std::vector<float> post_process() const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!target_) throw Error(VR_ERR_INVALID, "nothing rendered yet");
        float scale = samples_.first ? (float)samples_.second / (float)samples_.first : 0.0f;
        if (!std::isnormal(scale)) scale = 0.0f;
        const ColorManagementSettings& cm = settings_->color_management;
        return PostProcessingPass().render(*target_, PostProcessingData{scale, cm.gamma, cm.exposure, (int32_t)cm.tonemap});
    }
    const RenderTarget* target() const { return target_.get(); }

private:
    void one_shot() {
        try {
            {
                std::lock_guard<std::mutex> lock(mutex_);
                start_ = std::chrono::steady_clock::now();
                started_ = true;
                error_.clear();
            }
            auto accel = scene_->build_acceleration(ctx_);                                   // renderer.rs:58
            const RenderSettings rs = settings_->render;
            auto target = std::make_unique<RenderTarget>(accel, width_, height_, rs);        // clear, renderer.rs:55
            {
                std::lock_guard<std::mutex> lock(mutex_);
                target_ = std::move(target);
                samples_ = {0, rs.total_samples};
            }
            uint32_t samples = 0;
            const uint32_t total = rs.total_samples;
            const auto t0 = std::chrono::steady_clock::now();
            bool alive = iterative_render(*target_, *accel, rs, 1);                          // renderer.rs:67
            samples += 1;
            const double single = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            set_samples(samples, total);
            uint32_t per_frame = single > 0 ? (uint32_t)(rs.update_frequency / single) : total;
            per_frame = std::min(std::max(per_frame, 1u), std::max(total - samples, 1u));    // renderer.rs:78-81
            while (alive && samples < total && !cancel_) {
                // renderer.rs:86-91 passes samples_per_frame (overshooting on the last batch); this mirror draws
                // delta_samples so that exactly total_samples are accumulated
                const uint32_t delta = std::min(per_frame, total - samples);
                alive = iterative_render(*target_, *accel, rs, delta);
                if (alive) samples += delta;
                else samples = target_->stats().samples_done;
                set_samples(samples, total);
            }
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lock(mutex_);
            error_ = e.what();
        }
        std::lock_guard<std::mutex> lock(mutex_);
        end_ = std::chrono::steady_clock::now();
        rendering_ = false;
    }
    void set_samples(uint32_t done, uint32_t total) {
        std::lock_guard<std::mutex> lock(mutex_);
        samples_ = {done, total};
    }

    const Context& ctx_;
    std::shared_ptr<Scene> scene_;
    std::shared_ptr<Settings> settings_;
    uint32_t width_, height_;
    std::unique_ptr<RenderTarget> target_;
    std::thread thread_;
    mutable std::mutex mutex_;
    std::atomic<bool> rendering_{false}, cancel_{false};
    std::pair<uint32_t, uint32_t> samples_{0, 0};
    bool started_ = false;
    std::chrono::steady_clock::time_point start_, end_;
    std::string error_;
};

}  // namespace voidray
