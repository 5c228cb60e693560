"""-m gpu: size-independent properties at BASELINE.json's full resolutions (the oracle would take
minutes there): determinism, additivity of accumulation, sharding invariance, alpha = call count,
finiteness, clamp bound, and the closed-form furnace value."""
import numpy as np
import pytest

from voidray_b200 import scenes
from voidray_b200.render import RenderTarget
from voidray_b200.scene import Environments, Materials, RenderSettings

from util import F32

pytestmark = pytest.mark.gpu


def test_config2_full_resolution_properties(ctx):
    scene, st, (w, h) = scenes.config2_mossy_ground()        # 1920x1080, textured + normal-mapped, indoor HDRI
    assert (w, h) == (1920, 1080)
    accel = scene.build_acceleration(ctx)
    rs = RenderSettings(total_samples=8, max_bounces=8)
    a = RenderTarget(accel, (w, h), rs)
    a.accumulate(8)
    img = a.read()
    assert np.all(np.isfinite(img)) and np.all(img[..., 3] == 1.0)
    # firefly clamp: a pixel mean is bounded by max(clamp, peak environment radiance) = 15 (primary misses are unclamped)
    assert img[..., :3].max() <= 15.0 + 1e-3 and img[..., :3].min() >= 0.0
    b = RenderTarget(accel, (w, h), rs)
    b.accumulate(8)
    assert np.array_equal(img, b.read())                      # deterministic
    # additivity: 3 + 5 samples == the same sample set
    c = RenderTarget(accel, (w, h), rs)
    c.accumulate(3)
    c.accumulate(5)
    img_c = c.read()
    assert np.all(img_c[..., 3] == 2.0)
    assert np.abs(img_c[..., :3] - img[..., :3]).max() <= 4e-6
    st_ = a.stats()
    assert st_.camera_samples == w * h * 8 and w * h * 8 <= st_.ray_segments <= w * h * 8 * 8


def test_config5_4k_sharding_invariance(ctx):
    scene, st, (w, h) = scenes.config5_combined()             # 3840x2160
    assert (w, h) == (3840, 2160)
    accel = scene.build_acceleration(ctx)
    one = RenderTarget(accel, (w, h), RenderSettings(total_samples=4, max_bounces=8))
    one.accumulate(4)
    want = one.read()
    del one
    total = np.zeros_like(want)
    for rank in range(2):
        t = RenderTarget(accel, (w, h), RenderSettings(total_samples=4, max_bounces=8, sample_offset=2 * rank))
        t.accumulate(2)
        total += t.read()
        del t
    assert np.abs(total[..., :3] - want[..., :3]).max() <= 2e-6
    assert np.all(np.isfinite(want)) and want[..., :3].mean() > 0.01


def test_furnace_at_full_hd(ctx):
    # closed-form: convex-free check — with max_bounces = 1 every hit pixel is black and every miss is E,
    # with a white dielectric the whole 1920x1080 image equals E bit for bit
    scene, st, (w, h) = scenes.config1_mushroom(1920, 1080, dof=False)
    scene.environment = Environments.uniform((0.25, 0.5, 0.75))
    scene.materials[0] = Materials.dielectric(1.5)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), RenderSettings(total_samples=2, max_bounces=64))
    tgt.accumulate(2)
    img = tgt.read()
    bad = np.any(img[..., :3] != np.array([0.25, 0.5, 0.75], F32), axis=2)
    # only paths still bouncing inside the (non-convex, open) glass mesh after 64 bounces can differ, and
    # they can only be darker
    assert bad.mean() < 1e-2
    assert np.all(img[..., :3] <= np.array([0.25, 0.5, 0.75], F32))


# ---- converged-image gates at BASELINE.json's own sizes (north_star: "converged images must agree with the reference CPU
# render at equal spp within a stated RMSE / relative-MSE tolerance, and a fixed-seed per-pixel mean test must pass").
# Tolerances (the same table is in BASELINE.md):
#   fixed seed, per-pixel mean      max |GPU - oracle| <= 2e-4 absolute on every pixel and channel (values are O(1); the
#                                   only non-exact functions on the path are acos / atan2 of the HDRI lookup and the
#                                   30-degree normal test: libm ulps), and >= 30 % of the pixels bit-equal
#   independent seeds, equal spp    relMSE(GPU, oracle_b) <= 2 x relMSE(oracle_c, oracle_b) + 1e-4 and
#                                   RMSE(GPU, oracle_b)   <= 2 x RMSE(oracle_c, oracle_b)   + 1e-3, i.e. the GPU image is
#                                   no farther from an oracle render than two oracle renders are from each other
#                                   (relMSE = mean((a - b)^2 / (b^2 + 1e-2)) over RGB)
def test_config1_full_size_fixed_seed_per_pixel_mean(oracle, ctx):
    from util import rel_mse, rmse
    scene, st, (w, h) = scenes.config1_mushroom()          # 800 x 600, 64 spp, depth 8, thin lens, HDRI
    assert (w, h, st.render.total_samples, st.render.max_bounces) == (800, 600, 64, 8)
    rs = st.render
    osc = oracle.OracleScene(scene)
    ref, c_ref = osc.render(w, h, rs, 64)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), rs)
    tgt.accumulate(64)
    img = tgt.read()
    assert np.array_equal(img[..., 3], ref[..., 3])
    d = np.abs(img[..., :3] - ref[..., :3])
    assert d.max() <= 2e-4, f"max per-pixel mean difference {d.max():.3e}"
    assert np.mean(np.all(img == ref, axis=2)) > 0.3
    assert abs(tgt.stats().ray_segments - c_ref.segments) <= max(8, c_ref.segments // 20000)
    # independent seeds at equal spp
    rs_b = RenderSettings(total_samples=64, max_bounces=8, seed=222)
    rs_c = RenderSettings(total_samples=64, max_bounces=8, seed=333)
    ref_b, _ = osc.render(w, h, rs_b, 64)
    ref_c, _ = osc.render(w, h, rs_c, 64)
    assert rel_mse(img, ref_b) <= 2.0 * rel_mse(ref_c, ref_b) + 1e-4
    assert rmse(img, ref_b) <= 2.0 * rmse(ref_c, ref_b) + 1e-3


def test_config2_full_size_fixed_seed_per_pixel_mean(oracle, ctx):
    scene, st, (w, h) = scenes.config2_mossy_ground()      # 1920 x 1080, albedo + normal textures, HDRI; 16 of the 256 spp
    assert (w, h) == (1920, 1080)
    rs = RenderSettings(total_samples=16, max_bounces=8)
    ref, c_ref = oracle.OracleScene(scene).render(w, h, rs, 16)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), rs)
    tgt.accumulate(16)
    img = tgt.read()
    d = np.abs(img[..., :3] - ref[..., :3])
    assert d.max() <= 2e-4, f"max per-pixel mean difference {d.max():.3e}"
    assert np.mean(np.all(img == ref, axis=2)) > 0.3
    assert abs(tgt.stats().ray_segments - c_ref.segments) <= max(8, c_ref.segments // 20000)
