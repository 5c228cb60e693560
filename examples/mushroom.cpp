// examples/mushroom.cpp — the reference's mushroom scene (voidray_app/src/examples/mushroom.rs) built entirely from
// files, the way the reference builds it: Scene::add_image_texture(path), Scene::add_mesh_from_file(path),
// Environments::hdri(path) — every file decoded by the library's own loaders (csrc/image_io.cpp, scene_build.cpp).
//
//   ./examples/mushroom <asset_dir> <hdri file> [spp] [width] [height] [out.ppm]
//
// The reference repository does not ship mushroom_normal.jpg, mossy_ground_normal.jpg or studio.exr: the normal
// maps fall back to wood_normal.tif and the environment is whatever lat-long image is passed (the tests write the
// closed-form studio stand-in as an OpenEXR file). Prints a digest of the accumulation buffer;
// tests/test_cpp_host.py compares it bit for bit with the Python host's render of the same files.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../include/voidray.hpp"

using namespace voidray;

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <asset_dir> <hdri file> [spp] [width] [height] [out.ppm]\n", argv[0]);
        return 2;
    }
    const std::string dir = std::string(argv[1]) + "/";
    const uint32_t spp = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 16;
    const uint32_t W = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 1000, H = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 1000;  // mushroom.rs:7
    const char* out_path = argc > 6 ? argv[6] : nullptr;
    try {
        Context ctx(0);
        Scene scene = Scene::empty();
        Settings settings;

        const TextureHandle mushroom_albedo = scene.add_image_texture(dir + "mushroom_albedo.jpg", SampleType::Bilinear);
        const TextureHandle mushroom_normal = scene.add_image_texture(dir + "wood_normal.tif", SampleType::Bilinear);
        const MaterialHandle mushroom_mtl = scene.add_material(Materials::lambertian_texture(mushroom_albedo, mushroom_normal));
        const SurfaceHandle mushroom = scene.add_mesh_from_file(dir + "mushroom.obj");
        scene.add_object(mushroom_mtl, mushroom);

        const TextureHandle ground_albedo = scene.add_image_texture(dir + "mossy_ground_albedo.jpg", SampleType::Bilinear);
        const TextureHandle ground_normal = scene.add_image_texture(dir + "wood_normal.tif", SampleType::Bilinear);
        const MaterialHandle ground_mtl = scene.add_material(Materials::lambertian_texture(ground_albedo, ground_normal));
        const SurfaceHandle ground = scene.add_mesh_from_file(dir + "mossy_ground.obj");
        scene.add_object(ground_mtl, ground);

        scene.camera.eye = Vec3{0.2f, 2.8f, -10.5f};
        scene.camera.direction = Vec3{0.0f, -0.2f, 1.0f};
        scene.camera.fov = 0.17f;
        scene.camera.has_dof = true;
        scene.camera.aperture = 0.17f;
        scene.camera.focal_point = Vec3{0.06f, 2.14f, 0.18f};

        settings.color_management.gamma = 1.0f;
        settings.color_management.exposure = 1.0f;
        settings.color_management.tonemap = Tonemap::ACES;
        settings.render.total_samples = spp;

        scene.environment = Environments::hdri(argv[2]);

        auto accel = scene.build_acceleration(ctx);
        RenderTarget target(accel, W, H, settings.render);
        iterative_render(target, *accel, settings.render, spp);
        const std::vector<float> a = target.read();
        uint64_t d = 1469598103934665603ull;
        const unsigned char* b = (const unsigned char*)a.data();
        for (size_t i = 0; i < a.size() * 4; ++i) d = (d ^ b[i]) * 1099511628211ull;
        const vr_stats st = target.stats();
        std::printf("direct digest %016llx segments %llu\n", (unsigned long long)d, (unsigned long long)st.ray_segments);
        if (out_path) {
            const ColorManagementSettings& cm = settings.color_management;
            const std::vector<float> img = PostProcessingPass().render(target, PostProcessingData{1.0f, cm.gamma, cm.exposure, (int32_t)cm.tonemap});
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
