"""-m gpu: closest-hit parity of the CUDA path (through the C ABI) with the oracle.
Gate (BASELINE.json north_star): primary-ray hit triangle ids bit-exact, distances within 1e-5
relative. The implementation evaluates Möller–Trumbore in the reference's operation order, so the
distances are asserted bit-equal as well."""
import numpy as np
import pytest

from voidray_b200 import scenes
from voidray_b200.render import RenderTarget
from voidray_b200.scene import (Camera, Environments, Materials, MeshData, PixelMapping, RenderSettings, Scene,
                                Surfaces)

from util import (F32, MISS, build_case_scene, flat_case_rays, flat_split_cases, obj_scene, quad_obj, random_rays,
                  scene_bounds, single_mesh_scene)

pytestmark = pytest.mark.gpu


def check_primary(oracle, ctx, scene, rs, w, h, samples=(0,)):
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    tgt = RenderTarget(accel, (w, h), rs)
    for smp in samples:
        _, _, s_ref, p_ref, t_ref, _ = osc.trace_primary(w, h, rs, smp)
        s, p, t = tgt.trace_primary(smp)
        assert np.array_equal(s, s_ref), f"surface ids differ at {int((s != s_ref).sum())} pixels"
        assert np.array_equal(p, p_ref), f"triangle ids differ at {int((p != p_ref).sum())} pixels"
        hit = s_ref != MISS
        assert np.all(np.abs(t[hit] - t_ref[hit]) <= 1e-5 * t_ref[hit])
        assert np.array_equal(t, t_ref)
    return float((s_ref != MISS).mean())


def test_primary_config1_full_size(oracle, ctx):
    # configs[0] at its full 800x600, with the depth-of-field camera (lens + jitter drawn from the stream)
    scene, st, (w, h) = scenes.config1_mushroom()
    frac = check_primary(oracle, ctx, scene, st.render, w, h, samples=(0, 63))
    assert 0.2 < frac < 0.4


def test_primary_config1_square_reference_mapping(oracle, ctx):
    # the reference's own pixel mapping (iterative.rs:26,33) on a square 600x600 target
    scene, st, _ = scenes.config1_mushroom(600, 600, dof=False)
    st.render.pixel_mapping = PixelMapping.Reference
    check_primary(oracle, ctx, scene, st.render, 600, 600)


def test_primary_non_square_reference_mapping(oracle, ctx):
    # W > H with the reference mapping: sheared rows and u32 wrap-around, reproduced as is
    scene, st, _ = scenes.config1_mushroom(96, 64, dof=False)
    st.render.pixel_mapping = PixelMapping.Reference
    check_primary(oracle, ctx, scene, st.render, 96, 64)
    check_primary(oracle, ctx, scene, st.render, 64, 96)


def test_primary_two_surfaces_and_materials_scene(oracle, ctx):
    scene, st, _ = scenes.config5_combined(480, 270)
    assert check_primary(oracle, ctx, scene, st.render, 480, 270) > 0.7
    scene, st, _ = scenes.config3_materials(480, 270)
    check_primary(oracle, ctx, scene, st.render, 480, 270)
    scene, st, _ = scenes.config2_mossy_ground(480, 270)
    check_primary(oracle, ctx, scene, st.render, 480, 270)


@pytest.mark.parametrize("maker", [scenes.config2_mossy_ground, scenes.config3_materials, scenes.config5_combined])
def test_primary_full_baseline_resolution(oracle, ctx, maker):
    # the gate at BASELINE.json's own resolutions: 1920x1080 (configs 2, 3) and 3840x2160 (config 5)
    scene, st, (w, h) = maker()
    assert (w, h) in ((1920, 1080), (3840, 2160))
    check_primary(oracle, ctx, scene, st.render, w, h)


def test_primary_analytic_and_small_meshes(oracle, ctx):
    for fn in (scenes.example_cornell, scenes.example_spheres, scenes.example_material):
        scene, st, _ = fn()
        check_primary(oracle, ctx, scene, RenderSettings(total_samples=4, max_bounces=4), 200, 200, samples=(0, 2))


@pytest.mark.parametrize("name,n", [("cube.obj", 200000), ("mushroom.obj", 200000), ("mossy_ground.obj", 100000),
                                    ("material_testing_stand.obj", 100000), ("fancy_monkey.obj", 100000)])
def test_incoherent_rays(oracle, ctx, name, n):
    # secondary-ray-like workload: random origins and directions around each mesh
    scene = obj_scene(name)
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    o, d = random_rays(n, *scene_bounds(scene), seed=21)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
    s, p, t = accel.trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert 0.1 < (s_ref != MISS).mean()
    # rays that start on a surface (like every scattered ray) must not re-hit it below t = 1e-5
    hit = s_ref != MISS
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    o2 = (o[hit] + dn[hit] * t_ref[hit, None]).astype(F32)
    d2 = np.random.default_rng(5).normal(size=o2.shape).astype(F32)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o2, d2)
    s, p, t = accel.trace_rays(o2, d2)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)


def test_tie_rule_matches_reference_order(oracle, ctx):
    # coincident triangles: the reference returns the right-most leaf of its median-split tree
    # (bvh.rs:171); small meshes return the first index (mesh.rs:131); between surfaces the scene tree decides
    base = np.array([[-1, -1, 2], [1, -1, 2], [0, 1, 2]], F32)
    vs = [base] + [base + np.array([3.0 * k, 0, 0], F32) for k in range(1, 5)] + [base]
    big = MeshData.from_buffers(np.concatenate(vs), np.arange(18))
    small = MeshData.from_buffers(np.concatenate([base, base]), [0, 1, 2, 3, 4, 5])
    o = np.zeros((1, 3), F32)
    d = np.array([[0, 0, 1]], F32)
    for mesh in (big, small):
        scene = single_mesh_scene(mesh)
        osc = oracle.OracleScene(scene)
        accel = scene.build_acceleration(ctx)
        assert np.array_equal(accel.tie_ranks(0), osc.global_tie_rank(0))
        s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
        s, p, t = accel.trace_rays(o, d)
        assert (s[0], p[0], t[0]) == (s_ref[0], p_ref[0], t_ref[0])
    # two surfaces with coincident geometry + a sphere tangent at the same t
    scene = Scene.empty()
    m = scene.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    for mesh in (small, big, small):
        scene.add_object(m, scene.add_mesh(mesh))
    scene.add_object(m, scene.add_analytic_surface(Surfaces.sphere((0, 0, 3), 1.0)))
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
    s, p, t = accel.trace_rays(o, d)
    assert (s[0], p[0], t[0]) == (s_ref[0], p_ref[0], t_ref[0])
    for sf in range(3):
        assert np.array_equal(accel.tie_ranks(sf), osc.global_tie_rank(sf))


def test_tie_ranks_on_real_meshes(oracle, ctx):
    for name in ("mushroom.obj", "mossy_ground.obj"):
        scene = obj_scene(name)
        assert np.array_equal(scene.build_acceleration(ctx).tie_ranks(0), oracle.OracleScene(scene).global_tie_rank(0))


def test_million_triangle_field(oracle, ctx):
    # config 4's recipe at 16 x 15 copies = 1 067 520 triangles in ONE mesh: above the size where the commit sorts
    # the top of the reference-order tree and bins / partitions the top of the SAH tree on several threads
    scene, st, _ = scenes.config4_field(480, 270, 4, 8, nx=16, nz=15)
    assert scene.n_triangles() == 1067520
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    assert np.array_equal(accel.tie_ranks(0), osc.global_tie_rank(0))
    check_primary(oracle, ctx, scene, st.render, 480, 270)
    o, d = random_rays(100000, *scene_bounds(scene), seed=33)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
    s, p, t = accel.trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert 0.05 < (s_ref != MISS).mean()


def test_degenerate_inputs(oracle, ctx):
    # empty scene, zero-area triangles, axis-parallel rays, rays inside flat boxes
    empty = Scene.empty()
    accel = empty.build_acceleration(ctx)
    s, p, t = accel.trace_rays(np.zeros((4, 3), F32), np.eye(3, dtype=F32)[[0, 1, 2, 0]])
    assert np.all(s == MISS) and np.all(np.isinf(t))
    quad = Surfaces.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0))
    degenerate = MeshData.from_buffers(np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0], [1, 1, 0], [0, 2, 0]], F32),
                                       [0, 1, 2, 3, 4, 5])
    scene = Scene.empty()
    m = scene.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    scene.add_object(m, scene.add_mesh(quad))
    scene.add_object(m, scene.add_mesh(degenerate))
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    o = np.array([[0.25, 1, 0.5], [0.25, 1, 0.5], [0.5, 0.0, -1], [0.3, 0.3, 1], [0.5, 1, 0.5]], F32)
    d = np.array([[0, -1, 0], [0, 1, 0], [0, 0, 1], [0, 0, -1], [1, 0, 0]], F32)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
    s, p, t = accel.trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert s_ref[0] == 0 and t_ref[0] == F32(1.0)


# ---- scene-level culling of the reference (core/scene.rs:182-185 -> core/bvh.rs:132-160 over the surfaces) ----
# The Split boxes of the scene tree are unions of un-expanded Mesh::bounds(): a Split that is flat on an axis rejects
# every ray with a component along it (util/aabb.rs:109,126,143). The CUDA path reproduces that per ray.
@pytest.mark.parametrize("case", ["two_coplanar_quads", "three_coplanar_quads", "five_coplanar_quads",
                                  "flat_mesh_and_sphere", "flat_pair_inside_larger_scene", "two_coplanar_walls_x"])
def test_scene_level_split_culling_matches_reference(oracle, ctx, tmp_path, case):
    scene, _ = build_case_scene(flat_split_cases(tmp_path)[case])
    o, d = flat_case_rays(scene, 3)
    osc = oracle.OracleScene(scene)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d, oracle.MODE_FAITHFUL)
    s_all, _, _, _ = osc.trace_rays(o, d, oracle.MODE_BRUTE)
    s, p, t = scene.build_acceleration(ctx).trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    culled = (s_all != MISS) & (s_ref != s_all)
    if case == "flat_mesh_and_sphere":
        assert not culled.any() and (s_ref != MISS).mean() > 0.5
    else:
        assert culled.sum() > 100  # not vacuous: an un-culled closest hit finds what the reference drops


def test_scene_level_culling_four_judge_rays(oracle, ctx):
    # the case of VERDICT round 1: two coplanar quads (y = 0, x in [0, 1] and [2, 3]) as separate surfaces, four rays:
    # the reference misses all four
    scene = Scene.empty()
    m = scene.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    scene.add_object(m, scene.add_mesh(Surfaces.quad((0, 0, 0), (0, 0, 1), (1, 0, 1), (1, 0, 0))))
    scene.add_object(m, scene.add_mesh(Surfaces.quad((2, 0, 0), (2, 0, 1), (3, 0, 1), (3, 0, 0))))
    o = np.array([[0.5, 1, 0.5], [2.5, 1, 0.5], [0.5, 2, 0.25], [2.4, 1, 0.5]], F32)
    d = np.array([[0, -1, 0], [0.1, -1, 0.05], [0, -1, 0.125], [0.02, -1, 0.01]], F32)
    s_ref, p_ref, t_ref, _ = oracle.OracleScene(scene).trace_rays(o, d, oracle.MODE_FAITHFUL)
    assert np.all(s_ref == MISS) and np.all(np.isinf(t_ref))
    s, p, t = scene.build_acceleration(ctx).trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)


def test_scene_level_culling_beyond_the_mask_width(oracle, ctx, tmp_path):
    # 40 surfaces (> the 32 visibility bits a ray carries): winning candidates walk their ancestor chain instead
    surfaces = []
    for k in range(40):
        p = str(tmp_path / f"tile_{k}.obj")
        x0, z0, y = 2.0 * (k % 8), 2.0 * ((k // 8) % 4), float(k // 32)
        quad_obj(p, [(x0, y, z0), (x0, y, z0 + 1), (x0 + 1, y, z0 + 1), (x0 + 1, y, z0)])
        surfaces.append(("obj", p))
    scene, _ = build_case_scene(surfaces)
    o, d = flat_case_rays(scene, 9)
    osc = oracle.OracleScene(scene)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d, oracle.MODE_FAITHFUL)
    s_all, _, _, _ = osc.trace_rays(o, d, oracle.MODE_BRUTE)
    s, p, t = scene.build_acceleration(ctx).trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert ((s_all != MISS) & (s_ref != s_all)).sum() > 100 and (s_ref != MISS).sum() > 100


def tiled_floor_scene():
    """A floor tiled from four quads (an ordinary use of Surfaces::quad) under a sphere and an emitter: the scene tree
    has flat Splits over the tile pairs, so the wavefront's rays (primary and scattered) are culled as in the reference."""
    scene = Scene.empty()
    grey = scene.add_material(Materials.lambertian((0.6, 0.6, 0.6)))
    red = scene.add_material(Materials.lambertian((0.7, 0.2, 0.2)))
    for k, (x0, z0) in enumerate([(-2, -2), (0, -2), (-2, 0), (0, 0)]):
        scene.add_object(grey if k % 2 == 0 else red,
                         scene.add_mesh(Surfaces.quad((x0, 0, z0), (x0, 0, z0 + 2), (x0 + 2, 0, z0 + 2), (x0 + 2, 0, z0))))
    scene.add_object(scene.add_material(Materials.metal((0.8, 0.8, 0.8), 0.1)),
                     scene.add_analytic_surface(Surfaces.sphere((0.0, 0.75, 0.0), 0.75)))
    scene.add_object(scene.add_material(Materials.colored_emissive((1.0, 0.9, 0.8), 4.0)),
                     scene.add_mesh(Surfaces.quad((-1, 3, -1), (-1, 3, 1), (1, 3, 1), (1, 3, -1))))
    scene.camera = Camera.look_at((3.0, 2.5, 5.0), (0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 0.6)
    scene.environment = Environments.uniform((0.4, 0.5, 0.6))
    return scene


def test_scene_level_culling_in_the_wavefront(oracle, ctx):
    # k_trace (not the gate kernel): primary ids and the accumulated image of a tiled-floor scene against the
    # reference-faithful oracle, bit for bit (no libm on this path)
    scene = tiled_floor_scene()
    rs = RenderSettings(total_samples=8, max_bounces=6)
    w, h = 160, 120
    check_primary(oracle, ctx, scene, rs, w, h, samples=(0, 5))
    osc = oracle.OracleScene(scene)
    ref, _ = osc.render(w, h, rs, 8)
    brute, _ = osc.render(w, h, rs, 8, mode=oracle.MODE_BRUTE)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), rs)
    tgt.accumulate(8)
    img = tgt.read()
    assert np.abs(img - ref).max() <= 2e-6
    assert np.abs(brute - ref).max() > 0.05  # the culling is visible in the image: not a vacuous comparison
