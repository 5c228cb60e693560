// scene_file.h — test infrastructure: reads the flat scene file tests/scene_file.py writes into a HostScene, the way the
// vr_scene_* entry points of csrc/abi.cu fill it (same material conversion, same texture / environment records), and
// builds the DeviceScene over host memory that vr_scene_commit builds over device memory.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scene_build.h"

namespace vr_test {
using namespace vr;

struct Reader {
    FILE* f;
    bool ok = true;
    template <class T>
    T get() {
        T v{};
        if (std::fread(&v, sizeof(T), 1, f) != 1) ok = false;
        return v;
    }
    void floats(std::vector<float>& v, size_t n) {
        v.resize(n);
        if (n && std::fread(v.data(), 4, n, f) != n) ok = false;
    }
};

inline bool read_scene(const char* path, HostScene& sc) {
    Reader r{std::fopen(path, "rb")};
    if (!r.f) return false;
    if (r.get<uint32_t>() != 0x56525343u) return false;
    const uint32_t n_tex = r.get<uint32_t>();
    for (uint32_t i = 0; i < n_tex; ++i) {
        HostTexture t;
        t.w = r.get<uint32_t>();
        t.h = r.get<uint32_t>();
        t.sample_type = r.get<int32_t>();
        r.floats(t.rgb, (size_t)3 * t.w * t.h);
        pack_texture_rgba8(t);  // as vr_scene_add_texture_rgb32f does in the -DVR_TEX8 build
        sc.textures.push_back(std::move(t));
    }
    const uint32_t n_surf = r.get<uint32_t>();
    for (uint32_t i = 0; i < n_surf; ++i) {
        HostSurface sf;
        const uint32_t kind = r.get<uint32_t>();
        if (kind == 0) {  // vr_scene_add_mesh
            HostMesh m;
            m.n_vertices = r.get<uint32_t>();
            const uint32_t n_idx = r.get<uint32_t>();
            r.floats(m.pos, (size_t)3 * m.n_vertices);
            r.floats(m.uv, (size_t)2 * m.n_vertices);
            r.floats(m.nrm, (size_t)3 * m.n_vertices);
            m.idx.resize(n_idx);
            if (n_idx && std::fread(m.idx.data(), 4, n_idx, r.f) != n_idx) r.ok = false;
            m.idx.resize(n_idx - n_idx % 3);  // chunks_exact(3), mesh.rs:79
            sf.kind = 0;
            sf.mesh = (uint32_t)sc.meshes.size();
            sc.meshes.push_back(std::move(m));
        } else if (kind == 1) {
            sf.kind = 1;
            for (int a = 0; a < 3; ++a) sf.center[a] = r.get<float>();
            sf.radius_or_height = r.get<float>();
        } else {
            sf.kind = 2;
            sf.radius_or_height = r.get<float>();
        }
        sc.surfaces.push_back(sf);
    }
    const uint32_t n_mat = r.get<uint32_t>();
    for (uint32_t i = 0; i < n_mat; ++i) {  // vr_scene_add_material (csrc/abi.cu)
        MaterialRec m{};
        m.kind = r.get<int32_t>();
        for (int a = 0; a < 3; ++a) m.color[a] = r.get<float>();
        m.param = r.get<float>();
        const int32_t albedo = r.get<int32_t>(), normal = r.get<int32_t>();
        m.albedo_tex = m.kind == 0 ? albedo : -1;
        m.normal_tex = m.kind == 0 ? normal : -1;
        m.index = r.get<float>();
        m.roughness = r.get<float>();
        m.metallic = r.get<float>();
        m.emittance = r.get<float>();
        m.transparent = r.get<int32_t>() ? 1 : 0;
        if (m.kind == 3) {  // Emission::new: color * strength, simple.rs:168-172
            const float c[3] = {m.color[0], m.color[1], m.color[2]};
            for (int a = 0; a < 3; ++a) m.color[a] = c[a] * m.param;
        }
        sc.materials.push_back(m);
    }
    const uint32_t n_obj = r.get<uint32_t>();
    for (uint32_t i = 0; i < n_obj; ++i) {
        const uint32_t material = r.get<uint32_t>(), surface = r.get<uint32_t>();
        sc.objects.push_back(HostObject{surface, material});
    }
    for (int a = 0; a < 3; ++a) sc.camera.eye[a] = r.get<float>();
    for (int a = 0; a < 3; ++a) sc.camera.direction[a] = r.get<float>();
    for (int a = 0; a < 3; ++a) sc.camera.up[a] = r.get<float>();
    sc.camera.fov = r.get<float>();
    sc.camera.has_dof = r.get<int32_t>();
    sc.camera.aperture = r.get<float>();
    for (int a = 0; a < 3; ++a) sc.camera.focal_point[a] = r.get<float>();
    sc.env_kind = r.get<int32_t>();
    if (sc.env_kind == 1) {
        for (int a = 0; a < 3; ++a) sc.env_color[a] = r.get<float>();
    } else if (sc.env_kind == 2) {
        sc.env_image.w = r.get<uint32_t>();
        sc.env_image.h = r.get<uint32_t>();
        sc.env_image.sample_type = 1;
        r.floats(sc.env_image.rgb, (size_t)3 * sc.env_image.w * sc.env_image.h);
    }
    std::fclose(r.f);
    return r.ok;
}

// RGB f32 -> the RGBA f32 texel records the kernels fetch (k_expand_rgb)
struct HostTexels {
    std::vector<Quad> texels;
    TextureRec rec;
};
inline void expand_texture(const HostTexture& t, HostTexels& out) {
    const size_t n = (size_t)t.w * t.h;
    if (!t.rgba8.empty()) {  // upload_texture of the -DVR_TEX8 build: the RGBA8 form goes to the device
        out.rec.texels = t.rgba8.data();
        out.rec.width = t.w;
        out.rec.height = t.h;
        out.rec.sample_type = t.sample_type;
        out.rec.pad = 1;
        return;
    }
    out.texels.resize(n);
    for (size_t i = 0; i < n; ++i) out.texels[i] = Quad{t.rgb[3 * i], t.rgb[3 * i + 1], t.rgb[3 * i + 2], 0.0f};
    out.rec.texels = out.texels.data();
    out.rec.width = t.w;
    out.rec.height = t.h;
    out.rec.sample_type = t.sample_type;
    out.rec.pad = 0;
}

// what vr_scene_commit leaves in scene->dev, over host memory
struct HostDeviceScene {
    FlatScene flat;
    std::vector<HostTexels> textures;
    std::vector<TextureRec> texture_recs;
    HostTexels env;
    DeviceScene ds{};
    bool build(const HostScene& sc, std::string& err) {
        if (!flatten_scene(sc, flat, err)) return false;
        textures.resize(sc.textures.size());
        for (size_t i = 0; i < sc.textures.size(); ++i) {
            expand_texture(sc.textures[i], textures[i]);
            texture_recs.push_back(textures[i].rec);
        }
        ds.nodes = flat.nodes.data();
        ds.tri_isect = flat.tri_isect.data();
        ds.tri_shade = flat.tri_shade.data();
        ds.tri_surface = flat.tri_surface.data();
        ds.tri_prim = flat.tri_prim.data();
        ds.materials = sc.materials.data();
        ds.textures = texture_recs.data();
        ds.analytics = flat.analytics.data();
        ds.n_analytics = (uint32_t)flat.analytics.size();
        ds.n_tris = flat.n_tris;
        ds.scene_tree = flat.scene_tree.data();
        ds.surface_node = flat.surface_node.data();
        ds.n_scene_nodes = (uint32_t)flat.scene_tree.size();
        ds.n_surfaces = (uint32_t)flat.surface_node.size();
        for (const MaterialRec& m : sc.materials)
            if (m.kind == 5) ds.has_microfacet = 1;
        for (int a = 0; a < 3; ++a) {
            ds.grid_min[a] = flat.grid_min[a];
            ds.grid_extent[a] = flat.grid_extent[a];
            ds.env_color[a] = sc.env_color[a];
        }
        ds.env_kind = sc.env_kind;
        if (sc.env_kind == 2) {
            expand_texture(sc.env_image, env);
            ds.env_tex = env.rec;
        }
        ds.camera = flat.camera;
        return true;
    }
};
}  // namespace vr_test
