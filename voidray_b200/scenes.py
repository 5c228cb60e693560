"""Named scene recipes: the reference's example scenes (voidray_app/src/examples/*.rs) and the five
BASELINE.json configurations. Each returns (Scene, Settings, (width, height)) like the
reference's `examples::*::scene()`.

Substitutions for inputs missing from the reference checkout (.MISSING_LARGE_BLOBS), all stated in
DESIGN.md: studio.exr / indoor.exr -> assets.synth_hdri("studio" | "indoor"); mushroom_normal.jpg and
mossy_ground_normal.jpg -> wood_normal.tif; material_testing_main.obj -> fancy_monkey.obj.
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from .assets import asset_path, load_image_rgb32f, load_obj, merge_meshes, synth_hdri, transform_mesh
from .scene import (Camera, Environments, Materials, RenderSettings, SampleType, Scene, Settings, Surfaces, Tonemap,
                    hex_color)

F32 = np.float32


@lru_cache(maxsize=None)
def _mesh(name: str):
    return load_obj(asset_path(name))


@lru_cache(maxsize=None)
def _image(name: str):
    img = load_image_rgb32f(asset_path(name))
    img.setflags(write=False)
    return img


def _mushroom_camera(scene: Scene, dof: bool = True) -> None:
    # examples/mushroom.rs:27-30 — the default camera (Scene::empty) with eye/direction/fov/dof overwritten
    scene.camera.eye = (0.2, 2.8, -10.5)
    scene.camera.direction = (0.0, -0.2, 1.0)
    scene.camera.fov = 0.17
    scene.camera.dof = (0.17, (0.06, 2.14, 0.18)) if dof else None


def _aces_settings(total_samples: int, max_bounces: int, exposure: float = 1.0) -> Settings:
    s = Settings()
    s.render = RenderSettings(total_samples=total_samples, max_bounces=max_bounces)
    s.color_management.gamma = 1.0
    s.color_management.exposure = exposure
    s.color_management.tonemap = Tonemap.ACES
    return s


# ---- BASELINE.json configs ------------------------------------------------------------------------
def config1_mushroom(width: int = 800, height: int = 600, spp: int = 64, max_bounces: int = 8, dof: bool = True,
                     normal_map: bool = False):
    """configs[0]: mushroom.obj + studio HDRI, 800x600, 64 spp, max depth 8."""
    scene = Scene.empty()
    albedo = scene.add_image_texture(_image("mushroom_albedo.jpg"), SampleType.Bilinear)
    if normal_map:
        normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Bilinear)
        mtl = scene.add_material(Materials.lambertian_texture(albedo, normal))
    else:
        mtl = scene.add_material(Materials.lambertian_texture_no_normal(albedo))
    mushroom = scene.add_mesh(_mesh("mushroom.obj"))
    scene.add_object(mtl, mushroom)
    _mushroom_camera(scene, dof)
    scene.environment = Environments.hdri(synth_hdri("studio"))
    return scene, _aces_settings(spp, max_bounces), (width, height)


def config2_mossy_ground(width: int = 1920, height: int = 1080, spp: int = 256, max_bounces: int = 8):
    """configs[1]: mossy_ground.obj with albedo/normal textures + indoor HDRI, 1920x1080, 256 spp."""
    scene = Scene.empty()
    albedo = scene.add_image_texture(_image("mossy_ground_albedo.jpg"), SampleType.Bilinear)
    normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Bilinear)
    mtl = scene.add_material(Materials.lambertian_texture(albedo, normal))
    ground = scene.add_mesh(_mesh("mossy_ground.obj"))
    scene.add_object(mtl, ground)
    _mushroom_camera(scene, True)
    scene.environment = Environments.hdri(synth_hdri("indoor"))
    return scene, _aces_settings(spp, max_bounces), (width, height)


def config3_materials(width: int = 1920, height: int = 1080, spp: int = 1024, max_bounces: int = 8):
    """configs[2]: material_testing stand + four stand-in main meshes with diffuse / metal / dielectric /
    wood-textured materials, indoor HDRI, 1920x1080, 1024 spp."""
    scene = Scene.empty()
    stand = scene.add_mesh(_mesh("material_testing_stand.obj"))
    stand_mtl = scene.add_material(Materials.lambertian(hex_color(0x0F0F0F)))  # examples/material.rs:18-20
    scene.add_object(stand_mtl, stand)
    wood_albedo = scene.add_image_texture(_image("wood_albedo.tif"), SampleType.Bilinear)   # examples/spheres.rs:48-52
    wood_normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Nearest)
    mats = [
        Materials.lambertian(hex_color(0x7CA3E7)),
        Materials.metal(hex_color(0xE78999), 0.05),
        Materials.dielectric(1.5),
        Materials.lambertian_texture(wood_albedo, wood_normal),
    ]
    monkey = _mesh("fancy_monkey.obj")
    offsets = [(-1.0, 0.3, 0.4), (-0.55, 0.3, 1.0), (0.6, 0.3, 1.0), (1.05, 0.3, 0.4)]
    for mat, off in zip(mats, offsets):
        s = scene.add_mesh(transform_mesh(monkey, 0.0, off, scale=0.2))
        scene.add_object(scene.add_material(mat), s)
    scene.camera = Camera.look_at((0.5, 0.6, 5.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), float(F32(np.pi) / F32(6.0)))
    scene.environment = Environments.hdri(synth_hdri("indoor"))
    return scene, _aces_settings(spp, max_bounces, exposure=2.0), (width, height)


def _hash01(i: int, seed: int) -> float:
    x = (i * 0x9E3779B1 + seed * 0x85EBCA77) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x / 4294967296.0


def mushroom_field(nx: int, nz: int, spacing: float = 1.5, seed: int = 1234):
    """Baked world-space copies of mushroom.obj on an nx x nz grid, each rotated about y."""
    base = _mesh("mushroom.obj")
    copies = []
    for k in range(nx * nz):
        ix, iz = k % nx, k // nx
        off = ((ix - (nx - 1) / 2.0) * spacing, 0.0, (iz - (nz - 1) / 2.0) * spacing)
        copies.append(transform_mesh(base, 2.0 * np.pi * _hash01(k, seed), off))
    return merge_meshes(copies)


def config4_field(width: int = 1920, height: int = 1080, spp: int = 256, max_bounces: int = 8, nx: int = 48, nz: int = 47):
    """configs[3]: synthetic 10M-triangle mesh field (48 x 47 mushrooms = 10 034 688 triangles) + studio HDRI."""
    scene = Scene.empty()
    albedo = scene.add_image_texture(_image("mushroom_albedo.jpg"), SampleType.Bilinear)
    mtl = scene.add_material(Materials.lambertian_texture_no_normal(albedo))
    field = scene.add_mesh(mushroom_field(nx, nz))
    scene.add_object(mtl, field)
    scene.camera = Camera.look_at((0.0, 14.0, -0.62 * nz * 1.5 - 8.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 0.9)
    scene.environment = Environments.hdri(synth_hdri("studio"))
    return scene, _aces_settings(spp, max_bounces), (width, height)


def config5_combined(width: int = 3840, height: int = 2160, spp: int = 4096, max_bounces: int = 8):
    """configs[4]: mossy_ground + mushroom (examples/mushroom.rs verbatim, normal maps -> wood_normal.tif)."""
    scene = Scene.empty()
    mushroom_albedo = scene.add_image_texture(_image("mushroom_albedo.jpg"), SampleType.Bilinear)
    mushroom_normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Bilinear)
    mushroom_mtl = scene.add_material(Materials.lambertian_texture(mushroom_albedo, mushroom_normal))
    mushroom = scene.add_mesh(_mesh("mushroom.obj"))
    scene.add_object(mushroom_mtl, mushroom)
    ground_albedo = scene.add_image_texture(_image("mossy_ground_albedo.jpg"), SampleType.Bilinear)
    ground_normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Bilinear)
    ground_mtl = scene.add_material(Materials.lambertian_texture(ground_albedo, ground_normal))
    ground = scene.add_mesh(_mesh("mossy_ground.obj"))
    scene.add_object(ground_mtl, ground)
    _mushroom_camera(scene, True)
    scene.environment = Environments.hdri(synth_hdri("studio"))
    return scene, _aces_settings(spp, max_bounces), (width, height)


# ---- the reference's own example scenes -----------------------------------------------------------
def example_cornell():
    """voidray_app/src/examples/cornell.rs"""
    scene = Scene.empty()
    settings = Settings()
    settings.color_management.gamma = 1.0
    settings.color_management.exposure = 2.0
    settings.color_management.tonemap = Tonemap.Filmic
    red = scene.add_material(Materials.lambertian((0.65, 0.05, 0.05)))
    white = scene.add_material(Materials.lambertian((0.73, 0.73, 0.73)))
    green = scene.add_material(Materials.lambertian((0.12, 0.45, 0.15)))
    light = scene.add_material(Materials.emissive(15.0))
    floor = scene.add_mesh(Surfaces.quad((0, 0, 0), (0, 0, 555), (555, 0, 555), (555, 0, 0)))
    red_wall = scene.add_mesh(Surfaces.quad((0, 0, 0), (0, 0, 555), (0, 555, 555), (0, 555, 0)))
    green_wall = scene.add_mesh(Surfaces.quad((555, 0, 0), (555, 0, 555), (555, 555, 555), (555, 555, 0)))
    back_wall = scene.add_mesh(Surfaces.quad((0, 0, 555), (555, 0, 555), (555, 555, 555), (0, 555, 555)))
    ceil = scene.add_mesh(Surfaces.quad((0, 555, 0), (0, 555, 555), (555, 555, 555), (555, 555, 0)))
    light_plane = scene.add_mesh(Surfaces.quad((213, 554, 227), (213, 554, 332), (343, 554, 332), (343, 554, 227)))
    scene.add_object(white, floor)
    scene.add_object(green, green_wall)
    scene.add_object(red, red_wall)
    scene.add_object(white, back_wall)
    scene.add_object(white, ceil)
    scene.add_object(light, light_plane)
    sph = scene.add_analytic_surface(Surfaces.sphere((555.0 / 2.0, 100.0, 555.0 / 2.0), 100.0))
    glass = scene.add_material(Materials.dielectric(1.33))
    sph_inner = scene.add_analytic_surface(Surfaces.sphere((555.0 / 2.0, 100.0, 555.0 / 2.0), 99.9))
    glass_inner = scene.add_material(Materials.lambertian(hex_color(0x0F1BF0)))
    scene.add_object(glass, sph)
    scene.add_object(glass_inner, sph_inner)
    scene.camera = Camera.look_at((278.0, 278.0, -800.0), (278.0, 278.0, 0.0), (0.0, 1.0, 0.0),
                                  float(F32(40.0) * F32(np.pi) / F32(180.0)))
    return scene, settings, (500, 500)


def example_spheres():
    """voidray_app/src/examples/spheres.rs (indoor.exr -> synth_hdri("indoor"))"""
    scene = Scene.empty()
    settings = Settings()
    red = scene.add_material(Materials.metal(hex_color(0xE78999), 0.05))
    yellow = scene.add_material(Materials.dielectric(1.5))
    green = scene.add_material(Materials.metal(hex_color(0xB3E7AA), 0.1))
    blue = scene.add_material(Materials.metal(hex_color(0x7CA3E7), 0.01))
    grey = scene.add_material(Materials.metal(hex_color(0xAAAAAA), 0.01))
    diffuse = scene.add_material(Materials.lambertian(hex_color(0x7CA3E7)))
    light_mtl = scene.add_material(Materials.colored_emissive(hex_color(0xFF0F0F), 200.0))
    for pos, mtl in [((0.5, 1.0, 4.0), red), ((3.15, 1.5, -0.7), yellow), ((0.6, 0.6, -2.0), green),
                     ((-1.7, 1.1, -0.2), blue), ((1.2, 0.5, 0.4), grey)]:
        sph = scene.add_analytic_surface(Surfaces.sphere(pos, pos[1]))
        scene.add_object(mtl, sph)
    glass_2 = scene.add_material(Materials.dielectric(1.5))
    sph = scene.add_analytic_surface(Surfaces.sphere((1.5, 0.7, -3.1), 0.7))
    sph_inner = scene.add_analytic_surface(Surfaces.sphere((1.5, 0.695, -3.1), 0.695))
    scene.add_object(glass_2, sph)
    scene.add_object(diffuse, sph_inner)
    light = scene.add_analytic_surface(Surfaces.sphere((1.2, 8.0, -1.5), 2.0))
    scene.add_object(light_mtl, light)
    scene.camera = Camera.look_at((0.7166, 2.8803, -9.2992), (0.8673, 0.9557, 0.2095), (0.0, 1.0, 0.0), 0.6911)
    scene.camera.dof = (0.12, (0.1, 0.6, -2.0))
    saloon_albedo = scene.add_image_texture(_image("wood_albedo.tif"), SampleType.Bilinear)
    saloon_normal = scene.add_image_texture(_image("wood_normal.tif"), SampleType.Nearest)
    gnd = scene.add_analytic_surface(Surfaces.ground_plane(0.0))
    gnd_mat = scene.add_material(Materials.lambertian_texture(saloon_albedo, saloon_normal))
    scene.add_object(gnd_mat, gnd)
    scene.environment = Environments.hdri(synth_hdri("indoor"))
    settings.color_management.tonemap = Tonemap.ACES
    settings.color_management.gamma = 1.0
    settings.color_management.exposure = 2.0
    return scene, settings, (1000, 1000)


def example_material():
    """voidray_app/src/examples/material.rs"""
    scene = Scene.empty()
    settings = Settings()
    ground = scene.add_analytic_surface(Surfaces.ground_plane(0.0))
    uv_test = scene.add_image_texture(_image("uv_test.png"), SampleType.Nearest)
    ground_mat = scene.add_material(Materials.lambertian_texture_no_normal(uv_test))
    scene.add_object(ground_mat, ground)
    sphere = scene.add_analytic_surface(Surfaces.sphere((0.0, 0.5, 0.0), 0.5))
    sphere_mat = scene.add_material(Materials.lambertian_bsdf(hex_color(0xFF3030)))
    scene.add_object(sphere_mat, sphere)
    scene.camera.eye = (0.5, 0.6, 5.0)
    scene.environment = Environments.hdri(synth_hdri("indoor"))
    return scene, settings, (1000, 1000)


def example_mushroom():
    """voidray_app/src/examples/mushroom.rs at its own 1000x1000 (studio.exr / normal jpgs substituted)."""
    scene, settings, _ = config5_combined(1000, 1000, 100, 10)
    return scene, settings, (1000, 1000)


CONFIGS = {
    "config1_mushroom": config1_mushroom,
    "config2_mossy_ground": config2_mossy_ground,
    "config3_materials": config3_materials,
    "config4_field": config4_field,
    "config5_combined": config5_combined,
}
