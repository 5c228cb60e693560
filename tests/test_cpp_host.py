"""The C++ host mirror (include/voidray.hpp) and its example: compiles and links on CPU; on the GPU the example's
render of the reference's cornell scene is bit-identical to the Python host's (both drive the same library)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")


def _build(name="cornell"):
    subprocess.check_call(["make", "-C", EXAMPLES, "-s"])
    return os.path.join(EXAMPLES, name)


def _fnv1a(buf: np.ndarray) -> str:
    d = 1469598103934665603
    for b in np.ascontiguousarray(buf).tobytes():
        d = ((d ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{d:016x}"


def test_cpp_host_compiles_and_links():
    exe = _build()
    assert os.access(exe, os.X_OK)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libvoidray_cuda.so" in out and "not found" not in out.split("libvoidray_cuda.so")[1].splitlines()[0]


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_cpp_host_fails_loudly_without_gpu():
    r = subprocess.run([_build(), "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA" in r.stderr


@pytest.mark.gpu
def test_cpp_example_matches_python_host(ctx):
    from voidray_b200 import scenes
    from voidray_b200.render import RenderTarget
    from voidray_b200.scene import RenderSettings
    spp = 8
    r = subprocess.run([_build(), str(spp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"direct digest ([0-9a-f]{16}) segments (\d+)", r.stdout)
    assert m, r.stdout
    assert re.search(rf"samples {spp}/{spp}", r.stdout), r.stdout       # the progressive driver completed
    scene, settings, dims = scenes.example_cornell()
    tgt = RenderTarget(scene.build_acceleration(ctx), dims, RenderSettings(total_samples=spp, max_bounces=10))
    tgt.accumulate(spp)
    assert _fnv1a(tgt.read()) == m.group(1)
    assert tgt.stats().ray_segments == int(m.group(2))


@pytest.mark.gpu
def test_cpp_file_driven_mushroom_matches_python_host(ctx, tmp_path):
    """examples/mushroom.cpp builds the reference's mushroom scene from files only (add_image_texture(path),
    add_mesh_from_file(path), Environments::hdri(path)); the Python host does the same through its own path-taking
    calls. Both decode with the library's loaders, so the renders are bit-identical."""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2

    from voidray_b200 import assets
    from voidray_b200.render import RenderTarget
    from voidray_b200.scene import (Camera, Environments, Materials, RenderSettings, SampleType, Scene)
    asset_dir = os.path.join(ROOT, "assets")
    hdri = str(tmp_path / "studio.exr")
    assert cv2.imwrite(hdri, assets.synth_hdri("studio", 512, 256)[:, :, ::-1].astype(np.float32))
    spp, w, h = 4, 200, 160
    r = subprocess.run([_build("mushroom"), asset_dir, hdri, str(spp), str(w), str(h)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"direct digest ([0-9a-f]{16}) segments (\d+)", r.stdout)
    assert m, r.stdout

    scene = Scene.empty()
    p = lambda n: os.path.join(asset_dir, n)  # noqa: E731
    ma = scene.add_image_texture(p("mushroom_albedo.jpg"), SampleType.Bilinear)
    mn = scene.add_image_texture(p("wood_normal.tif"), SampleType.Bilinear)
    scene.add_object(scene.add_material(Materials.lambertian_texture(ma, mn)), scene.add_mesh_from_file(p("mushroom.obj")))
    ga = scene.add_image_texture(p("mossy_ground_albedo.jpg"), SampleType.Bilinear)
    gn = scene.add_image_texture(p("wood_normal.tif"), SampleType.Bilinear)
    scene.add_object(scene.add_material(Materials.lambertian_texture(ga, gn)), scene.add_mesh_from_file(p("mossy_ground.obj")))
    scene.camera = Camera(eye=(0.2, 2.8, -10.5), direction=(0.0, -0.2, 1.0), up=(0.0, 1.0, 0.0), fov=0.17,
                          dof=(0.17, (0.06, 2.14, 0.18)))
    scene.environment = Environments.hdri(hdri)
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), RenderSettings(total_samples=spp))
    tgt.accumulate(spp)
    assert _fnv1a(tgt.read()) == m.group(1)
    assert tgt.stats().ray_segments == int(m.group(2))
