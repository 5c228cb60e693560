"""Oracle self-checks for the integrator (core/tracer.rs:19-56, render/iterative.rs:11-55) against
closed-form radiance, since the reference holds no golden images."""
import numpy as np

from voidray_b200.scene import (Camera, Environments, Materials, PixelMapping, RenderMode, RenderSettings, Scene,
                                Surfaces)

from util import F32


def sphere_scene(material, env=(0.5, 0.5, 0.5)):
    s = Scene.empty()
    sp = s.add_analytic_surface(Surfaces.sphere((0, 0, 0), 1.0))
    s.add_object(s.add_material(material), sp)
    s.camera = Camera.look_at((0, 0, 6), (0, 0, 0), (0, 1, 0), 0.5)
    s.environment = Environments.uniform(env) if env is not None else None
    return s


W = H = 32
CENTER = [H // 2 * W + W // 2, (H // 2 - 1) * W + W // 2 - 1]   # pixels well inside the sphere
CORNER = [0, W - 1, W * H - 1]                                    # pixels that miss it


def render(oracle, scene, **kw):
    rs = RenderSettings(total_samples=4, max_bounces=8, **kw)
    acc, _ = oracle.OracleScene(scene).render(W, H, rs, 4, n_threads=2)
    return acc.reshape(-1, 4)


def test_furnace_convex_lambertian(oracle):
    # a convex Lambertian body in a uniform environment E: hit pixels = min(rho * E, clamp), misses = E
    img = render(oracle, sphere_scene(Materials.lambertian((0.6, 0.4, 0.2))))
    for p in CENTER:
        assert np.array_equal(img[p, :3], np.array([0.6, 0.4, 0.2], F32) * F32(0.5))
        assert img[p, 3] == 1.0
    for p in CORNER:
        assert np.array_equal(img[p, :3], np.array([0.5, 0.5, 0.5], F32))


def test_firefly_clamp_per_level(oracle):
    # tracer.rs:50 / color.rs:30-36: min(att * L, 3) per channel; the miss itself is not clamped
    img = render(oracle, sphere_scene(Materials.lambertian((0.5, 0.2, 0.1)), env=(10.0, 10.0, 10.0)))
    for p in CENTER:
        assert np.array_equal(img[p, :3], np.array([3.0, 2.0, 1.0], F32))
    for p in CORNER:
        assert np.array_equal(img[p, :3], np.array([10.0, 10.0, 10.0], F32))
    img = render(oracle, sphere_scene(Materials.lambertian((0.5, 0.2, 0.1)), env=(10.0, 10.0, 10.0)), firefly_clamp=1.5)
    assert np.array_equal(img[CENTER[0], :3], np.array([1.5, 1.5, 1.0], F32))


def test_max_bounces_terminates_black(oracle):
    # depth == max_bounces returns BLACK (tracer.rs:28): with one bounce a hit contributes att * 0
    rs = RenderSettings(total_samples=4, max_bounces=1)
    acc, c = oracle.OracleScene(sphere_scene(Materials.lambertian((0.6, 0.4, 0.2)))).render(W, H, rs, 4, n_threads=2)
    img = acc.reshape(-1, 4)
    assert np.array_equal(img[CENTER[0], :3], np.zeros(3, F32))
    assert np.array_equal(img[CORNER[0], :3], np.array([0.5, 0.5, 0.5], F32))
    assert c.segments == W * H * 4   # exactly one scene.hit per camera sample
    rs0 = RenderSettings(total_samples=4, max_bounces=0)
    acc0, c0 = oracle.OracleScene(sphere_scene(Materials.lambertian((0.6, 0.4, 0.2)))).render(W, H, rs0, 4, n_threads=1)
    assert not acc0[..., :3].any() and c0.segments == 0


def test_emission_and_no_environment(oracle):
    img = render(oracle, sphere_scene(Materials.colored_emissive((1.0, 0.5, 0.25), 2.0), env=None))
    assert np.array_equal(img[CENTER[0], :3], np.array([2.0, 1.0, 0.5], F32))
    assert np.array_equal(img[CORNER[0], :3], np.zeros(3, F32))     # environment None -> BLACK (tracer.rs:31)
    img = render(oracle, sphere_scene(Materials.emissive(15.0), env=None))
    assert np.array_equal(img[CENTER[0], :3], np.array([3.0, 3.0, 3.0], F32))  # clamped at its own level


def test_dielectric_conserves_uniform_environment(oracle):
    # attenuation (1,1,1) at every bounce: every path that escapes returns exactly E
    rs = RenderSettings(total_samples=4, max_bounces=64)
    acc, _ = oracle.OracleScene(sphere_scene(Materials.dielectric(1.5), env=(0.25, 0.5, 0.75))).render(W, H, rs, 4, n_threads=2)
    img = acc.reshape(-1, 4)
    for p in CENTER + CORNER:
        assert np.array_equal(img[p, :3], np.array([0.25, 0.5, 0.75], F32))


def test_schlick_values(oracle):
    assert abs(oracle.schlick(1.0, 1.5) - 0.04) < 1e-7          # normal incidence: ((1-n)/(1+n))^2
    assert abs(oracle.schlick(0.0, 1.5) - 1.0) < 1e-7           # grazing
    assert abs(oracle.schlick(0.5, 1.0 / 1.5) - (0.04 + 0.96 * 0.5 ** 5)) < 1e-6


def test_metal_mirror_reflects_environment(oracle):
    # fuzz 0: one bounce into the uniform environment -> albedo * E
    img = render(oracle, sphere_scene(Materials.metal((0.8, 0.6, 0.4), 0.0)))
    assert np.array_equal(img[CENTER[0], :3], np.array([0.8, 0.6, 0.4], F32) * F32(0.5))


def test_lambertian_bsdf_quirk(oracle):
    # blanket BSDF impl (traits.rs:23-40) with LambertianBSDF's pdf = 1.0 and uniform-sphere wi:
    # attenuation = rho/pi * |wi.n|, and half the directions continue *into* the sphere
    img = render(oracle, sphere_scene(Materials.lambertian_bsdf((0.9, 0.9, 0.9))))
    v = img[CENTER[0], :3]
    assert np.all(v > 0.0) and np.all(v < 0.9 / np.pi * 0.5 + 1e-6)


def test_normal_render_mode(oracle):
    img = render(oracle, sphere_scene(Materials.lambertian((0.6, 0.4, 0.2))), render_mode=RenderMode.Normal)
    n = img[CENTER[0], :3] * 2 - 1          # 0.5 * n + 0.5 (tracer.rs:40)
    assert abs(np.linalg.norm(n) - 1.0) < 1e-3 and n[2] > 0.95
    assert np.array_equal(img[CORNER[0], :3], np.array([0.5, 0.5, 0.5], F32))


def test_iterative_render_accumulation_semantics(oracle):
    # iterative.rs:45-51: each call adds sum(samples)/total_samples and 1.0 to alpha
    scene = sphere_scene(Materials.lambertian((0.6, 0.4, 0.2)))
    osc = oracle.OracleScene(scene)
    rs = RenderSettings(total_samples=8, max_bounces=4)
    acc, _ = osc.render(W, H, rs, 4, n_threads=1)
    assert np.all(acc[..., 3] == 1.0)
    assert np.array_equal(acc.reshape(-1, 4)[CORNER[0], :3], np.array([0.25, 0.25, 0.25], F32))
    acc, _ = osc.render(W, H, rs, 4, accum=acc, sample_offset=4, n_threads=1)
    assert np.all(acc[..., 3] == 2.0)
    assert np.array_equal(acc.reshape(-1, 4)[CORNER[0], :3], np.array([0.5, 0.5, 0.5], F32))
    # the thread count does not change the result (each pixel's stream is keyed by (pixel, sample))
    a1, _ = osc.render(W, H, rs, 8, n_threads=1)
    a4, _ = osc.render(W, H, rs, 8, n_threads=4)
    assert np.array_equal(a1, a4)


def test_pixel_mapping_modes(oracle):
    scene = sphere_scene(Materials.lambertian((0.6, 0.4, 0.2)))
    scene.camera = Camera.look_at((0.4, 0.3, 6), (0, 0, 0), (0, 1, 0), 0.5)
    osc = oracle.OracleScene(scene)
    fixed = RenderSettings(total_samples=2, max_bounces=3, pixel_mapping=PixelMapping.Fixed)
    ref = RenderSettings(total_samples=2, max_bounces=3, pixel_mapping=PixelMapping.Reference)
    # square: the reference's y = index / height is the intended mapping
    a, _ = osc.render(24, 24, fixed, 2, n_threads=1)
    b, _ = osc.render(24, 24, ref, 2, n_threads=1)
    assert np.array_equal(a, b)
    # non-square (iterative.rs:26 divides by the height): the image is sheared
    a, _ = osc.render(32, 24, fixed, 2, n_threads=1)
    b, _ = osc.render(32, 24, ref, 2, n_threads=1)
    assert not np.array_equal(a, b)
    # fixed mapping: the sphere is centred horizontally up to the camera offset, image row 0 is the top
    hit = np.any(a[..., :3] != F32(0.5), axis=2)
    rows = np.where(hit.any(1))[0]
    assert rows.min() > 0 and rows.max() < 23
