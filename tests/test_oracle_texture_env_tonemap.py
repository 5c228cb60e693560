"""Oracle self-checks: texture lookups (core/texture.rs:52-98), HDRI lookup
(voidray_common/src/environments.rs:57-86), tonemap curves (shaders/tonemapping.glsl)."""
import numpy as np

from voidray_b200.scene import Environments, Materials, SampleType, Scene, Surfaces

from util import F32

TW, TH = 4, 3


def tex_scene(sample_type):
    s = Scene.empty()
    img = np.arange(TW * TH * 3, dtype=F32).reshape(TH, TW, 3)   # texel (row, col) = 12*row + 3*col + c
    t = s.add_image_texture(img, sample_type)
    sp = s.add_analytic_surface(Surfaces.sphere((0, 0, 0), 1.0))
    s.add_object(s.add_material(Materials.lambertian_texture_no_normal(t)), sp)
    return s, img


def test_nearest_sampling_edges(oracle):
    s, img = tex_scene(SampleType.Nearest)
    osc = oracle.OracleScene(s)
    uv = np.array([[0.0, 0.999], [1.0, 0.999], [-0.25, 0.999], [0.26, 0.0], [0.26, 1.0], [2.26, -0.4]], F32)
    got = osc.texture_sample(0, uv)
    assert np.array_equal(got[0], img[0, 0])       # x = 0, y = 0.003
    assert np.array_equal(got[1], img[0, 0])       # u % 1 == 0: u = 1 wraps to column 0
    assert np.array_equal(got[2], img[0, 3])       # u < 0: u -= trunc(u) - 1 -> 0.75 -> column 3
    assert np.array_equal(got[3], img[2, 1])       # v = 0 -> y = H, clamped to the last row
    assert np.array_equal(got[4], img[2, 1])       # v = 1 -> 1 - (1 % 1) = 1 -> same
    assert np.array_equal(got[5], img[1, 1])       # v = -0.4 -> 0.6 -> y = 1.2


def test_bilinear_flat_index_bleed(oracle):
    s, img = tex_scene(SampleType.Bilinear)
    osc = oracle.OracleScene(s)
    flat = img.reshape(-1, 3)
    # interior: plain bilinear between (row 1, col 1..2) and (row 2, col 1..2)
    u, v = 0.375, 0.5     # x = 1.5, y = 1.5
    got = osc.texture_sample(0, np.array([[u, v]], F32))[0]
    want = 0.5 * (0.5 * img[1, 1] + 0.5 * img[1, 2]) + 0.5 * (0.5 * img[2, 1] + 0.5 * img[2, 2])
    assert np.allclose(got, want, rtol=1e-6)
    # right edge: x0 = W-1, the "+1" tap is the first texel of the next row (texture.rs:67-73)
    u, v = 0.9375, 0.5    # x = 3.75, y = 1.5
    got = osc.texture_sample(0, np.array([[u, v]], F32))[0]
    top = 0.25 * flat[1 * TW + 3] + 0.75 * flat[1 * TW + 4]
    bot = 0.25 * flat[2 * TW + 3] + 0.75 * flat[(2 * TW + 4) % (TW * TH)]
    assert np.allclose(got, 0.5 * top + 0.5 * bot, rtol=1e-6)
    # v = 0: y = H, y0 = H-1, ay = 1 -> the whole weight is on row H, which wraps to row 0
    got = osc.texture_sample(0, np.array([[0.125, 0.0]], F32))[0]   # x = 0.5
    assert np.allclose(got, 0.5 * img[0, 0] + 0.5 * img[0, 1], rtol=1e-6)


def hdri_scene(w=16, h=8, seed=5):
    s = Scene.empty()
    img = np.random.default_rng(seed).uniform(0, 4, (h, w, 3)).astype(F32)
    s.environment = Environments.hdri(img)
    return s, img


def ref_hdri(img, d):
    """environments.rs:80-86 + :57-76 in float64 (tolerance comparison)."""
    h, w, _ = img.shape
    flat = img.reshape(-1, 3).astype(np.float64)
    out = []
    for v in d:
        v = v / np.linalg.norm(v)
        theta = np.arccos(np.clip(-v[1], -1, 1))
        phi = np.arctan2(-v[2], v[0]) + np.pi
        x = phi / (2 * np.pi) * w
        y = (h - 1) - theta / np.pi * h
        x0 = min(int(x) if x > 0 else 0, w - 1)
        y0 = min(int(y) if y > 0 else 0, h - 1)
        ax, ay = x - x0, y - y0
        n = w * h
        t = lambda i: flat[i % n]  # noqa: E731
        top = t(y0 * w + x0) * (1 - ax) + t(y0 * w + x0 + 1) * ax
        bot = t((y0 + 1) * w + x0) * (1 - ax) + t((y0 + 1) * w + x0 + 1) * ax
        out.append(top * (1 - ay) + bot * ay)
    return np.array(out)


def test_hdri_lookup_matches_formula(oracle):
    s, img = hdri_scene()
    osc = oracle.OracleScene(s)
    d = np.random.default_rng(2).normal(size=(500, 3)).astype(F32)
    got = osc.environment_sample(d)
    want = ref_hdri(img, d.astype(np.float64))
    assert np.allclose(got, want, rtol=2e-4, atol=2e-4)


def test_hdri_poles_and_seam(oracle):
    s, img = hdri_scene()
    osc = oracle.OracleScene(s)
    h, w, _ = img.shape
    # straight up: theta = pi -> y = (H-1) - H = -1 -> `as usize` saturates to 0, ay = -1 (extrapolation)
    up = osc.environment_sample(np.array([[0, 1, 0]], F32))[0]
    want = ref_hdri(img, np.array([[0.0, 1.0, 0.0]]))[0]
    assert np.allclose(up, want, rtol=1e-4, atol=1e-4)
    # straight down: theta = 0 -> y = H-1 exactly, ay = 0 -> last row
    down = osc.environment_sample(np.array([[0, -1, 0]], F32))[0]
    assert np.allclose(down, ref_hdri(img, np.array([[0.0, -1.0, 0.0]]))[0], rtol=1e-4, atol=1e-4)
    assert np.all(np.isfinite(up)) and np.all(np.isfinite(down))


def tonemap_ref(c, scale, gamma, exposure, mode):
    c = np.asarray(c, np.float64) * scale * 2.0 ** exposure
    if mode == 1:
        m_in = np.array([[0.59719, 0.076, 0.0284], [0.35458, 0.90834, 0.13383], [0.04823, 0.01566, 0.83777]]).T
        m_out = np.array([[1.60475, -0.10208, -0.00327], [-0.53108, 1.10813, -0.07276], [-0.07367, -0.00605, 1.07602]]).T
        c = m_in @ c
        c = (c * (c + 0.0245786) - 0.000090537) / (c * (0.983729 * c + 0.432951) + 0.238081)
        c = m_out @ c
    elif mode == 2:
        luma = c @ np.array([0.2126, 0.7152, 0.0722])
        c = c * (luma * (1 + luma / 4.0) / (1 + luma) / luma)
    elif mode == 3:
        c = np.maximum(0, c - 0.004)
        c = (c * (6.2 * c + 0.5)) / (c * (6.2 * c + 1.7) + 0.06)
    elif mode == 4:
        A, B, C, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
        f = lambda x: ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F  # noqa: E731
        c = f(c * 2.0) / f(11.2)
    with np.errstate(invalid="ignore"):
        return c if gamma == 1.0 else c ** (1.0 / gamma)


def test_tonemap_spot_values(oracle):
    vals = [(0.0, 0.0, 0.0), (0.18, 0.18, 0.18), (1.0, 1.0, 1.0), (16.0, 8.0, 0.5), (0.3, 0.6, 0.1)]
    acc = np.zeros((1, len(vals), 4), F32)
    acc[0, :, :3] = vals
    for mode in range(5):
        for scale, gamma, exposure in ((1.0, 1.0, 0.0), (2.0, 2.2, 1.0)):
            out = oracle.resolve(acc, scale, gamma, exposure, mode)
            assert np.all(out[..., 3] == 1.0)
            for k, v in enumerate(vals):
                if mode == 2 and v == (0.0, 0.0, 0.0):
                    assert np.all(np.isnan(out[0, k, :3]))   # Reinhard divides by luma (tonemapping.glsl:14)
                    continue
                want = tonemap_ref(v, scale, gamma, exposure, mode)
                got = out[0, k, :3]
                # Uncharted2's `- E/F` cancels in f32 (abs error ~1.6e-5 before the division by white = 0.067)
                atol = 5e-4 if mode == 4 else 2e-6
                ok = np.isclose(got, want, rtol=2e-5, atol=atol) | (np.isnan(got) & np.isnan(want))
                assert np.all(ok), (mode, v, got, want)
    # anchors: Uncharted2 maps the white point to 1, Filmic(0) = 0, None is the identity at gamma 1
    w = np.zeros((1, 1, 4), F32)
    w[0, 0, :3] = 11.2 / 2.0
    assert np.allclose(oracle.resolve(w, 1.0, 1.0, 0.0, 4)[0, 0, :3], 1.0, rtol=1e-6)
    assert np.array_equal(oracle.resolve(acc, 1.0, 1.0, 0.0, 3)[0, 0, :3], np.zeros(3, F32))
    assert np.array_equal(oracle.resolve(acc, 1.0, 1.0, 0.0, 0)[0, :, :3], acc[0, :, :3])
