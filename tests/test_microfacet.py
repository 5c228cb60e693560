"""MicrofacetBSDF (voidray_common/src/microfacet.rs) through the blanket BSDFMaterial impl
(core/traits.rs:23-40): oracle sanity on CPU, CUDA-vs-oracle parity on the GPU.
exp / ln / atan / sin / cos differ by a few ulp between glibc and CUDA, so per-sample radiance is compared
with an absolute tolerance instead of bit-equality."""
import numpy as np
import pytest

from voidray_b200 import MicrofacetBSDF, hex_color
from voidray_b200.scene import RenderSettings

from test_oracle_shading import CENTER, CORNER, sphere_scene
from util import F32

W = H = 32
MATERIALS = {
    "diffuse": MicrofacetBSDF.diffuse((0.8, 0.6, 0.4)),
    "specular": MicrofacetBSDF.specular((0.8, 0.6, 0.4), 0.3),
    "metallic": MicrofacetBSDF.metallic(hex_color(0xE7B959), 0.2),
    "clear": MicrofacetBSDF.clear(1.5, 0.1),
    "transparent": MicrofacetBSDF.transparent((0.9, 0.5, 0.5), 1.33, 0.4),
    "light": MicrofacetBSDF.light((1.0, 1.0, 1.0), 5.0),
}


def test_constructors_match_reference():
    m = MicrofacetBSDF.clear(1.5, 0.1)
    assert (m.kind, m.color, m.index, m.roughness, m.metallic, m.transparent) == (5, (1.0, 1.0, 1.0), 1.5, 0.1, 0.0, True)
    m = MicrofacetBSDF.metallic((0.1, 0.2, 0.3), 0.25)
    assert (m.index, m.roughness, m.metallic, m.transparent) == (1.5, 0.25, 1.0, False)
    m = MicrofacetBSDF.light((1, 1, 1), 7.0)
    assert (m.index, m.roughness, m.emittance) == (1.0, 1.0, 7.0)


@pytest.mark.parametrize("name", sorted(MATERIALS))
def test_oracle_microfacet_is_sane(oracle, name):
    scene = sphere_scene(MATERIALS[name], env=(0.5, 0.5, 0.5))
    rs = RenderSettings(total_samples=32, max_bounces=8)
    acc, c = oracle.OracleScene(scene).render(W, H, rs, 32, n_threads=2)
    img = acc.reshape(-1, 4)
    assert np.all(np.isfinite(img[:, :3])) and np.all(img[:, :3] >= 0.0) and np.all(img[:, :3] <= 3.0 + 1e-6)
    assert np.array_equal(img[CORNER[0], :3], np.array([0.5, 0.5, 0.5], F32))     # misses are untouched
    assert c.segments > W * H * 32
    if MATERIALS[name].transparent:
        assert img[CENTER[0], :3].max() > 0.1
    else:
        # reference quirk: the blanket impl passes the *incoming* direction as `wo` (traits.rs:30) and
        # HitRecord::new makes the normal face the ray, so n.wo < 0 on every hit and the opaque branch of
        # bsdf() (microfacet.rs:128-131) returns BLACK: opaque MicrofacetBSDF objects render black.
        assert np.array_equal(img[CENTER[0], :3], np.zeros(3, F32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MATERIALS))
def test_cuda_microfacet_matches_oracle(oracle, ctx, name):
    from voidray_b200.render import RenderTarget
    scene = sphere_scene(MATERIALS[name], env=(0.5, 0.5, 0.5))
    rs = RenderSettings(total_samples=16, max_bounces=8)
    w = h = 48
    osc = oracle.OracleScene(scene)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), rs)
    rng = np.random.default_rng(12)
    px = rng.integers(0, w * h, 20000).astype(np.uint32)
    sm = rng.integers(0, 16, 20000).astype(np.uint32)
    ref = osc.sample_radiance(w, h, rs, px, sm)
    got = tgt.sample_radiance(px, sm)
    fin = np.isfinite(ref).all(axis=1) & np.isfinite(got).all(axis=1)
    assert fin.mean() > 0.999
    err = np.abs(got[fin] - ref[fin]).max(axis=1)
    # transcendental ulps move attenuations by ~1e-6 relative; a flipped branch (Bernoulli / sign tests) moves a
    # whole path: allow 1 sample in 1000 beyond 1e-3
    assert (err > 1e-3).mean() <= 1e-3, f"{int((err > 1e-3).sum())} samples differ"
    assert np.median(err) <= 1e-5
    ref_img, _ = osc.render(w, h, rs, 16)
    tgt.accumulate(16)
    img = tgt.read()
    ok = np.isfinite(ref_img[..., :3]).all(axis=2) & np.isfinite(img[..., :3]).all(axis=2)
    assert ok.mean() > 0.999
    assert abs(img[..., :3][ok].mean() - ref_img[..., :3][ok].mean()) < 1e-3
