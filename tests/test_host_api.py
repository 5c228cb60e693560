"""Host-side mirror of the reference's scene API (voidray_b200/scene.py, assets.py): handle semantics,
constructors, loaders. No device needed."""
import numpy as np
import pytest

from voidray_b200 import (Camera, Environments, Materials, MeshData, RenderSettings, SampleType, Scene, Settings,
                          Surfaces, Tonemap, hex_color)
from voidray_b200.assets import asset_path, load_image_rgb32f, load_obj, synth_hdri
from voidray_b200 import scenes
from voidray_b200.distributed import shard_samples

F32 = np.float32


def test_handles_are_sequential_indices():
    s = Scene.empty()
    assert s.add_material(Materials.lambertian((1, 0, 0))) == 0
    assert s.add_material(Materials.metal((1, 1, 1), 0.1)) == 1
    assert s.add_analytic_surface(Surfaces.sphere((0, 0, 0), 1)) == 0
    assert s.add_mesh(Surfaces.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0))) == 1
    assert s.add_object(1, 0) == 0
    assert s.add_image_texture(np.zeros((2, 2, 3), F32), SampleType.Nearest) == 0
    assert s.objects[0].material == 1 and s.objects[0].surface == 0


def test_defaults_match_reference_settings():
    st = Settings()
    assert (st.render.total_samples, st.render.max_bounces, st.render.firefly_clamp) == (100, 10, 3.0)
    assert st.render.update_frequency == pytest.approx(0.1)
    assert st.color_management.tonemap == Tonemap.NONE and st.color_management.gamma == pytest.approx(2.2)
    assert [int(t) for t in Tonemap] == [0, 1, 2, 3, 4]      # Tonemap::as_i32, settings.rs:45-55
    # Scene::empty() camera: look_at((1,0,10), 0, +y, pi/6)
    cam = Scene.empty().camera
    assert cam.eye == (1.0, 0.0, 10.0) and cam.dof is None
    assert cam.fov == pytest.approx(np.pi / 6, rel=1e-6)


def test_look_at_matches_oracle(oracle):
    import ctypes as C
    lib = oracle.load()
    eye, center, up = (0.7166, 2.8803, -9.2992), (0.8673, 0.9557, 0.2095), (0.0, 1.0, 0.0)
    d = (C.c_float * 3)()
    u = (C.c_float * 3)()
    lib.vo_look_at((C.c_float * 3)(*eye), (C.c_float * 3)(*center), (C.c_float * 3)(*up), d, u)
    cam = Camera.look_at(eye, center, up, 0.6911)
    assert tuple(F32(x) for x in cam.direction) == tuple(F32(x) for x in d)
    assert tuple(F32(x) for x in cam.up) == tuple(F32(x) for x in u)
    assert abs(np.dot(cam.direction, cam.up)) < 1e-6


def test_hex_color_and_materials():
    assert hex_color(0xFF8000) == (1.0, float(F32(128) / F32(255)), 0.0)
    m = Materials.colored_emissive((1.0, 0.5, 0.25), 200.0)
    assert (m.kind, m.param) == (3, 200.0)
    m = Materials.lambertian_texture(2, 5)
    assert (m.kind, m.albedo_tex, m.normal_tex) == (0, 2, 5)
    assert Materials.lambertian_texture_no_normal(1).normal_tex == -1
    assert Materials.dielectric(1.5).param == 1.5
    q = Surfaces.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0))
    assert list(q.indices) == [0, 1, 2, 2, 0, 3] and not q.normals.any() and not q.uvs.any()


def test_obj_loader_dedups_index_triples_in_first_seen_order():
    cube = load_obj(asset_path("cube.obj"))
    assert cube.n_triangles == 12
    # every face vertex is a (v, vt, vn) triple; identical triples share one vertex
    assert cube.positions.shape[0] == len({tuple(r) for r in np.hstack([cube.positions, cube.uvs, cube.normals])})
    assert cube.indices[0] == 0 and cube.indices[1] == 1 and cube.indices[2] == 2   # first-seen order
    assert np.array_equal(cube.positions[0], np.array([-1, 1, -1], F32))            # "f 5/1/1 ..."
    assert np.array_equal(cube.uvs[0], np.array([0.875, 0.5], F32))
    mush = load_obj(asset_path("mushroom.obj"))
    assert mush.n_triangles == 4448 and mush.indices.max() == mush.positions.shape[0] - 1
    assert load_obj(asset_path("mossy_ground.obj")).n_triangles == 14699


def test_image_loader_matches_to_rgb32f():
    img = load_image_rgb32f(asset_path("test.png"))           # 2x2 RGBA -> RGB, /255
    assert img.shape == (2, 2, 3) and img.dtype == F32
    assert np.all((img * 255 == np.round(img * 255)))
    tif = load_image_rgb32f(asset_path("wood_normal.tif"))
    assert tif.shape == (512, 512, 3) and 0.0 <= tif.min() and tif.max() <= 1.0


def test_synthetic_hdris_are_deterministic():
    import hashlib
    a = synth_hdri("studio")
    assert a.shape == (1024, 2048, 3) and a.dtype == F32
    assert 19.0 < a.max() <= 21.0 and a.min() >= 0.0          # the 20x soft box exceeds firefly_clamp = 3
    b = synth_hdri("indoor")
    assert 14.0 < b.max() <= 16.5
    # bit-identical across machines: only IEEE +,-,*,/ in f64
    want = {"studio": None, "indoor": None}
    for name, img in (("studio", a), ("indoor", b)):
        want[name] = hashlib.sha256(img.tobytes()).hexdigest()
    import json, os
    golden = os.path.join(os.path.dirname(__file__), "golden", "synth_hdri_sha256.json")
    assert json.load(open(golden)) == want


def test_scene_recipes_build():
    for name, fn in scenes.CONFIGS.items():
        if name == "config4_field":
            scene, settings, dims = fn(nx=2, nz=2)
            assert scene.n_triangles() == 4 * 4448
        else:
            scene, settings, dims = fn()
        assert len(scene.objects) >= len(scene.surfaces)
        assert settings.render.max_bounces == 8 and settings.color_management.tonemap == Tonemap.ACES
    assert scenes.config1_mushroom()[2] == (800, 600)
    assert scenes.config5_combined()[2] == (3840, 2160) and scenes.config5_combined()[1].render.total_samples == 4096
    for fn in (scenes.example_cornell, scenes.example_spheres, scenes.example_material):
        scene, settings, dims = fn()
        assert dims[0] == dims[1]
    # the cornell recipe reproduces the reference's object order (green wall object at the red wall's index)
    cornell = scenes.example_cornell()[0]
    assert cornell.objects[1].surface == 2 and cornell.objects[2].surface == 1


def test_shard_samples_partitions_the_range():
    for total in (1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            ranges = [shard_samples(total, world, r) for r in range(world)]
            assert sum(c for _, c in ranges) == total
            pos = 0
            for off, cnt in ranges:
                assert off == pos
                pos += cnt
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    with pytest.raises(ValueError):
        shard_samples(8, 2, 2)
