// examples/cornell.cpp — the reference's cornell scene (voidray_app/src/examples/cornell.rs) built and rendered from
// C++ through include/voidray.hpp, the compiled-language mirror of the reference's host API.
//
//   make -C examples && ./examples/cornell [spp] [out.ppm]
//
// Prints a digest of the accumulation buffer (tests/test_cpp_host.py compares it with the Python host's render of
// the same scene: both drive the same library, so the buffers are bit-identical).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../include/voidray.hpp"

using namespace voidray;

static Scene cornell_scene(Settings& settings) {
    Scene scene = Scene::empty();
    settings.color_management.gamma = 1.0f;
    settings.color_management.exposure = 2.0f;
    settings.color_management.tonemap = Tonemap::Filmic;

    const MaterialHandle red = scene.add_material(Materials::lambertian(Color{0.65f, 0.05f, 0.05f}));
    const MaterialHandle white = scene.add_material(Materials::lambertian(Color{0.73f, 0.73f, 0.73f}));
    const MaterialHandle green = scene.add_material(Materials::lambertian(Color{0.12f, 0.45f, 0.15f}));
    const MaterialHandle light = scene.add_material(Materials::emissive(15.0f));

    const SurfaceHandle floor = scene.add_mesh(Surfaces::quad({0, 0, 0}, {0, 0, 555}, {555, 0, 555}, {555, 0, 0}));
    const SurfaceHandle red_wall = scene.add_mesh(Surfaces::quad({0, 0, 0}, {0, 0, 555}, {0, 555, 555}, {0, 555, 0}));
    const SurfaceHandle green_wall = scene.add_mesh(Surfaces::quad({555, 0, 0}, {555, 0, 555}, {555, 555, 555}, {555, 555, 0}));
    const SurfaceHandle back_wall = scene.add_mesh(Surfaces::quad({0, 0, 555}, {555, 0, 555}, {555, 555, 555}, {0, 555, 555}));
    const SurfaceHandle ceil = scene.add_mesh(Surfaces::quad({0, 555, 0}, {0, 555, 555}, {555, 555, 555}, {555, 555, 0}));
    const SurfaceHandle light_plane = scene.add_mesh(Surfaces::quad({213, 554, 227}, {213, 554, 332}, {343, 554, 332}, {343, 554, 227}));

    scene.add_object(white, floor);
    scene.add_object(green, green_wall);
    scene.add_object(red, red_wall);
    scene.add_object(white, back_wall);
    scene.add_object(white, ceil);
    scene.add_object(light, light_plane);

    const SurfaceHandle sph = scene.add_analytic_surface(Surfaces::sphere({555.0f / 2.0f, 100.0f, 555.0f / 2.0f}, 100.0f));
    const MaterialHandle glass = scene.add_material(Materials::dielectric(1.33f));
    const SurfaceHandle sph_inner = scene.add_analytic_surface(Surfaces::sphere({555.0f / 2.0f, 100.0f, 555.0f / 2.0f}, 99.9f));
    const MaterialHandle glass_inner = scene.add_material(Materials::lambertian(hex_color(0x0F1BF0)));
    scene.add_object(glass, sph);
    scene.add_object(glass_inner, sph_inner);

    // degrees_to_radians(40.0) = 40 * PI / 180 in f32 (util/math.rs:32-34)
    scene.camera = Camera::look_at({278.0f, 278.0f, -800.0f}, {278.0f, 278.0f, 0.0f}, {0.0f, 1.0f, 0.0f},
                                   40.0f * 3.14159265358979323846f / 180.0f);
    return scene;
}

int main(int argc, char** argv) {
    const uint32_t spp = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 64;
    const char* out_path = argc > 2 ? argv[2] : nullptr;
    const uint32_t W = 500, H = 500;  // examples/cornell.rs:13
    try {
        Context ctx(0);
        auto settings = std::make_shared<Settings>();
        auto scene = std::make_shared<Scene>(cornell_scene(*settings));
        settings->render.total_samples = spp;
        settings->render.max_bounces = 10;

        // (1) one blocking iterative_render call: deterministic, compared bit for bit with the Python host
        {
            auto accel = scene->build_acceleration(ctx);
            RenderTarget target(accel, W, H, settings->render);
            iterative_render(target, *accel, settings->render, spp);
            const std::vector<float> a = target.read();
            uint64_t d = 1469598103934665603ull;
            const unsigned char* b = (const unsigned char*)a.data();
            for (size_t i = 0; i < a.size() * 4; ++i) d = (d ^ b[i]) * 1099511628211ull;
            std::printf("direct digest %016llx segments %llu\n", (unsigned long long)d, (unsigned long long)target.stats().ray_segments);
        }

        // (2) the progressive driver, RenderThread::one_shot: 1-spp probe, batches sized by update_frequency
        Renderer renderer(ctx, scene, settings, W, H);
        renderer.execute(RenderAction::Render);
        renderer.join();
        const auto done = renderer.samples();
        const vr_stats st = renderer.target()->stats();
        std::printf("samples %u/%u  %.3f s  %.1f Msamples/s  %.1f Mrays/s\n", done.first, done.second, renderer.elapsed_time(),
                    (double)st.camera_samples / st.seconds / 1e6, (double)st.ray_segments / st.seconds / 1e6);

        const std::vector<float> accum = renderer.target()->read();
        double sum = 0.0;
        uint64_t digest = 1469598103934665603ull;  // FNV-1a over the buffer's bytes
        const unsigned char* bytes = (const unsigned char*)accum.data();
        for (size_t i = 0; i < accum.size() * 4; ++i) digest = (digest ^ bytes[i]) * 1099511628211ull;
        for (size_t i = 0; i < accum.size(); i += 4) sum += accum[i] + accum[i + 1] + accum[i + 2];
        std::printf("accum mean %.6f digest %016llx\n", sum / (3.0 * W * H), (unsigned long long)digest);

        if (out_path) {
            const std::vector<float> img = renderer.post_process();
            FILE* f = std::fopen(out_path, "wb");
            if (!f) throw Error(VR_ERR_INVALID, std::string("cannot write ") + out_path);
            std::fprintf(f, "P6\n%u %u\n255\n", W, H);
            for (size_t i = 0; i < (size_t)W * H; ++i)
                for (int c = 0; c < 3; ++c) {
                    float v = img[4 * i + c];
                    v = v != v ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v));
                    std::fputc((int)(v * 255.0f + 0.5f), f);
                }
            std::fclose(f);
        }
    } catch (const Error& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
