"""-m gpu: unit parity of the device-side building blocks with the oracle: generator, samplers,
texture / environment lookups, resolve (tonemap)."""
import ctypes as C

import numpy as np
import pytest

from voidray_b200 import _lib
from voidray_b200.render import RenderTarget
from voidray_b200.scene import Environments, Materials, RenderSettings, SampleType, Scene, Surfaces
from voidray_b200.assets import synth_hdri

from test_oracle_texture_env_tonemap import hdri_scene, tex_scene
from util import F32

pytestmark = pytest.mark.gpu


def test_generator_and_samplers_bit_exact(oracle, ctx):
    lib = _lib.load()
    for seed, pixel, sample in ((0x5EED0001, 0, 0), (0xDEADBEEFCAFEF00D, 479999, 4095), (1, 2 ** 31, 7)):
        out = np.empty(257, np.uint32)
        _lib.check(lib.vr_debug_rng_draws(ctx.handle, seed, pixel, sample, out.size, _lib.uptr(out)))
        assert np.array_equal(out, oracle.rng_draws(seed, pixel, sample, out.size))
        sp = np.empty((500, 3), F32)
        _lib.check(lib.vr_debug_unit_sphere(ctx.handle, seed, pixel, sample, 500, _lib.fptr(sp)))
        assert np.array_equal(sp, oracle.unit_sphere(seed, pixel, sample, 500))


@pytest.mark.parametrize("sample_type", [SampleType.Nearest, SampleType.Bilinear])
def test_texture_lookup_bit_exact(oracle, ctx, sample_type):
    s, img = tex_scene(sample_type)
    osc = oracle.OracleScene(s)
    accel = s.build_acceleration(ctx)
    rng = np.random.default_rng(4)
    uv = np.concatenate([
        rng.uniform(-3, 3, (20000, 2)),
        np.array([[0, 0], [1, 1], [0, 1], [1, 0], [-0.25, 0.999], [0.9375, 0.5], [0.125, 0.0], [-1.0, -1.0],
                  [2.0, -2.0], [1e-8, 1 - 1e-8], [-1e-8, -1e-8], [1e6 + 0.5, -1e6 - 0.5]]),
    ]).astype(F32)
    assert np.array_equal(accel.texture_sample(0, uv), osc.texture_sample(0, uv))


def test_texture_lookup_real_texture(oracle, ctx):
    from voidray_b200 import scenes
    scene, _, _ = scenes.config1_mushroom(64, 48)
    uv = np.random.default_rng(6).uniform(-1, 2, (50000, 2)).astype(F32)
    assert np.array_equal(scene.build_acceleration(ctx).texture_sample(0, uv),
                          oracle.OracleScene(scene).texture_sample(0, uv))


def test_environment_lookup(oracle, ctx):
    # acosf / atan2f are the only non-exact functions on the path: tolerance 2e-4 * max radiance
    for scene in (hdri_scene()[0], hdri_scene(64, 32, 9)[0]):
        d = np.random.default_rng(8).normal(size=(50000, 3)).astype(F32)
        d = np.concatenate([d, np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1],
                                         [-1, 0, 1e-9], [-1, 0, -1e-9]], F32)])
        ref = oracle.OracleScene(scene).environment_sample(d)
        got = scene.build_acceleration(ctx).environment_sample(d)
        err = np.abs(got - ref).max(axis=1)
        # the phi = 0 / 2 pi seam flips texel column for directions within an ulp of it
        assert (err > 1e-3).mean() < 1e-3 and np.median(err) < 1e-6
    s = Scene.empty()
    s.environment = Environments.hdri(synth_hdri("studio"))
    d = np.random.default_rng(9).normal(size=(100000, 3)).astype(F32)
    ref = oracle.OracleScene(s).environment_sample(d)
    got = s.build_acceleration(ctx).environment_sample(d)
    assert np.abs(got - ref).max() < 20.0 * 2e-4
    assert np.mean(np.all(got == ref, axis=1)) > 0.5
    s.environment = Environments.uniform((0.1, 0.2, 0.3))
    assert np.array_equal(s.build_acceleration(ctx).environment_sample(d[:10]), np.tile(np.array([0.1, 0.2, 0.3], F32), (10, 1)))


@pytest.mark.parametrize("dims", [(1920, 1080), (1031, 57), (8, 8)])
def test_resolve_all_tonemaps_any_size(oracle, ctx, dims):
    # post_process.glsl over the WHOLE target (the reference dispatch stops at 1024x1024, post_process.rs:74)
    w, h = dims
    s = Scene.empty()
    s.add_object(s.add_material(Materials.lambertian((0.5, 0.5, 0.5))), s.add_analytic_surface(Surfaces.sphere((0, 0, 0), 1)))
    s.environment = Environments.hdri(synth_hdri("studio"))
    rs = RenderSettings(total_samples=1, max_bounces=3)
    tgt = RenderTarget(s.build_acceleration(ctx), (w, h), rs)
    tgt.accumulate(1)
    acc = tgt.read()
    for mode in range(5):
        for scale, gamma, exposure in ((2.0, 1.0, 1.0), (2.0, 2.2, -0.5)):
            got = tgt.resolve(scale, gamma, exposure, mode)
            want = oracle.resolve(acc, scale, gamma, exposure, mode)
            assert got.shape == (h, w, 4) and np.all(got[..., 3] == 1.0)
            both_nan = np.isnan(got) & np.isnan(want)
            # f32 tonemap curves: 1e-5 relative + 1e-6 absolute (powf differs by an ulp or two)
            assert np.all(np.isclose(got, want, rtol=1e-5, atol=1e-6) | both_nan), (mode, gamma)
