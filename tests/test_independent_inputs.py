"""The oracle and the product are fed by the same host loaders (voidray_b200/assets.py), so a loader mistake would be
invisible to every CUDA-vs-oracle gate. These tests read every asset the BASELINE scenes use a second, independent way
(a minimal OBJ reader written here, plain PIL / OpenCV decodes) and require the very same arrays; and they restate the
core of the closest hit (Triangle::hit, core/mesh.rs:144-189) in numpy as a third opinion next to the oracle's C++ and
the kernels' CUDA. Parity stays "unpinned" (no output of the Rust reference exists here), but it is not single-source."""
import os

import numpy as np
import pytest

from voidray_b200 import assets
from voidray_b200.assets import asset_path

from util import F32, MISS, obj_scene, random_rays, scene_bounds

OBJS = ["cube.obj", "mushroom.obj", "fancy_monkey.obj", "mossy_ground.obj", "material_testing_stand.obj"]


def read_obj_minimal(path):
    """obj-rs 0.7.0 load_obj::<TexturedVertex, u32> (core/mesh.rs:46-74), written without looking at assets.load_obj:
    collect the statement lists, then walk the faces assigning vertex numbers to v/vt/vn triples as they first appear."""
    v, vt, vn, faces = [], [], [], []
    for raw in open(path):
        tok = raw.split("#")[0].split()
        if not tok:
            continue
        if tok[0] == "v":
            v.append([float(x) for x in tok[1:4]])
        elif tok[0] == "vt":
            vt.append([float(x) for x in (tok[1:3] + ["0"])[:2]])
        elif tok[0] == "vn":
            vn.append([float(x) for x in tok[1:4]])
        elif tok[0] == "f":
            assert len(tok) == 4, "triangulated faces only"
            faces.append([tuple(int(x) for x in corner.split("/")) for corner in tok[1:]])
    number, triples, indices = {}, [], []
    for face in faces:
        for corner in face:
            if corner not in number:
                number[corner] = len(triples)
                triples.append(corner)
            indices.append(number[corner])
    t = np.array(triples, dtype=np.int64) - 1  # the assets use positive 1-based indices only
    return (np.array(v, np.float64).astype(F32)[t[:, 0]], np.array(vt, np.float64).astype(F32)[t[:, 1]],
            np.array(vn, np.float64).astype(F32)[t[:, 2]], np.array(indices, np.uint32))


@pytest.mark.parametrize("name", OBJS)
def test_obj_assets_read_a_second_way(name):
    pos, uv, nrm, idx = read_obj_minimal(asset_path(name))
    for m in (assets.load_obj(asset_path(name)), assets.load_obj_native(asset_path(name))):  # host recipe loader, library loader
        assert np.array_equal(m.indices, idx)
        assert np.array_equal(m.positions, pos) and np.array_equal(m.uvs, uv) and np.array_equal(m.normals, nrm)


@pytest.mark.parametrize("name", ["mushroom_albedo.jpg", "mossy_ground_albedo.jpg", "wood_albedo.tif", "wood_normal.tif",
                                  "uv_test.png", "test.png"])
def test_image_assets_decoded_a_second_way(name):
    # image 0.24.3 `open(path).to_rgb32f()`: channel / 255 (8 bit) or / 65535 (16 bit), alpha dropped, no sRGB decode
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    raw = cv2.imread(asset_path(name), cv2.IMREAD_UNCHANGED)
    assert raw is not None
    if raw.ndim == 2:
        raw = np.repeat(raw[:, :, None], 3, axis=2)
    rgb = raw[:, :, 2::-1]  # BGR(A) -> RGB
    denom = {np.dtype(np.uint8): 255.0, np.dtype(np.uint16): 65535.0}[rgb.dtype]
    want = (rgb.astype(np.float64) / denom).astype(F32)  # the correctly rounded quotient, via f64
    host = assets.load_image_rgb32f(asset_path(name))   # PIL path: what the scene recipes feed the oracle and the library
    native = assets.load_image_native(asset_path(name))  # the library's own decoder
    assert host.shape == want.shape == native.shape
    if name.endswith(".jpg"):
        # no two JPEG decoders are bit-identical (IDCT rounding): PIL and OpenCV both wrap libjpeg-turbo, so they agree
        # exactly; the library's own decoder is held to +-3 levels on < 1 % of the samples (tests/test_image_io.py)
        assert np.array_equal(host, want)
        assert np.abs(native - want).max() <= 3.01 / 255.0 and np.mean(native != want) < 0.01
    else:
        assert np.array_equal(host, want) and np.array_equal(native, want)


def moller_trumbore_numpy(o, d, v0, v1, v2):
    """Triangle::hit, core/mesh.rs:144-189, one ray against every triangle, f32 operation by operation."""
    e1, e2 = v1 - v0, v2 - v0
    def cross(a, b):
        return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                         a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)
    def dot(a, b):
        return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]
    eps = F32(0.00001)
    h = cross(np.broadcast_to(d, e2.shape), e2)
    a = dot(e1, h)
    ok = ~((a > -eps) & (a < eps))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        f = F32(1.0) / a
        s = o - v0
        u = f * dot(s, h)
        ok &= ~((u < 0) | (u > 1))
        q = cross(s, e1)
        v = f * dot(np.broadcast_to(d, q.shape), q)
        ok &= ~((v < 0) | (u + v > 1))
        t = f * dot(e2, q)
    ok &= t > eps
    return np.where(ok, t, F32(np.inf))


@pytest.mark.parametrize("name,n", [("cube.obj", 2000), ("mushroom.obj", 1500), ("fancy_monkey.obj", 1500)])
def test_closest_hit_third_opinion_numpy(oracle, name, n):
    # brute force over every triangle in numpy f32 (cgmath 0.18 dot = (x*x + y*y) + z*z, cross as written above): the
    # oracle's reference-faithful traversal must report the same distance bit for bit, and the same triangle wherever
    # the minimum is unique
    scene = obj_scene(name)
    mesh = scene.surfaces[0]
    tri = mesh.indices.reshape(-1, 3)
    v0, v1, v2 = (mesh.positions[tri[:, k]].astype(F32) for k in range(3))
    o, d = random_rays(n, *scene_bounds(scene), seed=77)
    dn = d.astype(F32)
    # Ray::new normalises the direction (util/ray.rs:12-17); cgmath 0.18 InnerSpace::normalize(v) = v * (1 / v.magnitude()),
    # magnitude = sqrt(dot(v, v)), all in f32
    mag = np.sqrt(((dn[:, 0] * dn[:, 0] + dn[:, 1] * dn[:, 1]) + dn[:, 2] * dn[:, 2]).astype(F32)).astype(F32)
    dn = (dn * (F32(1.0) / mag)[:, None]).astype(F32)
    s_ref, p_ref, t_ref, _ = oracle.OracleScene(scene).trace_rays(o, d)
    hits = 0
    for i in range(n):
        t = moller_trumbore_numpy(o[i].astype(F32), dn[i], v0, v1, v2)
        best = t.min()
        assert best.view(np.uint32) == t_ref[i].view(np.uint32), (i, best, t_ref[i])
        if np.isfinite(best):
            hits += 1
            winners = np.flatnonzero(t == best)
            assert p_ref[i] in winners
            if len(winners) == 1:
                assert p_ref[i] == winners[0] and s_ref[i] == 0
        else:
            assert s_ref[i] == MISS
    assert hits > n // 10
