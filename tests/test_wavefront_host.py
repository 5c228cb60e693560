"""The CUDA kernels themselves on the CPU: tests/c/wavefront_host.cpp compiles voidray_b200/csrc/kernels.cu (k_raygen,
k_trace, k_shade, k_accumulate as nvcc sees them) through tests/c/host_shim.h — one OS thread per CUDA thread, warp
intrinsics over a per-warp barrier, real atomics — and runs one wavefront batch of a small frame. Checked here:
every wavefront hit equals the single-ray traversal (inside the harness), and the accumulated image equals the oracle's
render of the same scene with the same seed. This is the radiance gate of tests/test_gpu_radiance.py without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Camera, Environments, Materials, RenderSettings, Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"default": []}
EYE, CENTER, FOV = (0.2, 2.8, -10.5), (0.2, 0.8, -0.5), 0.17  # the mushroom example's view (examples/mushroom.rs:27-30)
ENV, ALBEDO = (0.75, 0.5, 0.25), (0.5, 0.625, 0.75)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    built = {}

    def get(variant):
        if variant not in built:
            exe = str(tmp_path_factory.mktemp("wf") / f"wavefront_host_{variant}")
            csrc = os.path.join(ROOT, "voidray_b200", "csrc")
            r = subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-DVR_HOST_SHIM", "-DVR_HOST_SIMT",
                                *VARIANTS[variant], "-I", os.path.join(ROOT, "tests", "c"), "-I", csrc, "-x", "c++",
                                os.path.join(csrc, "scene_build.cpp"), os.path.join(ROOT, "tests", "c", "wavefront_host.cpp"),
                                "-o", exe], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            built[variant] = exe
        return built[variant]
    return get


def oracle_render(oracle, name, w, h, spp, bounces, seed):
    scene = Scene.empty()
    scene.add_object(scene.add_material(Materials.lambertian(ALBEDO)), scene.add_mesh(load_obj(asset_path(name))))
    scene.environment = Environments.uniform(ENV)
    scene.camera = Camera.look_at(EYE, CENTER, (0, 1, 0), FOV)
    rs = RenderSettings(total_samples=spp, max_bounces=bounces, firefly_clamp=3.0, seed=seed)
    img, counters = oracle.OracleScene(scene).render(w, h, rs, spp)
    return img, counters


def test_kernels_on_the_cpu_match_the_oracle(oracle, harness, tmp_path, variant="default"):
    name, w, h, spp, bounces, seed = "mushroom.obj", 64, 48, 4, 6, 0x5EED0001
    out = str(tmp_path / "accum.bin")
    args = [harness(variant), asset_path(name), str(w), str(h), str(spp), str(bounces), hex(seed),
            *[repr(float(x)) for x in EYE], *[repr(float(x)) for x in CENTER], repr(FOV),
            *[repr(float(x)) for x in ENV], *[repr(float(x)) for x in ALBEDO], out]
    r = subprocess.run(args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 of " in r.stdout and "wavefront hits differ" in r.stdout
    img = np.fromfile(out, np.float32).reshape(h, w, 4)
    ref, counters = oracle_render(oracle, name, w, h, spp, bounces, seed)
    # the same number of scene.hit calls as the reference recursion makes, and at least a third of the camera rays hit
    segments = int(r.stdout.split(" segments")[0].split()[-1])
    assert segments == counters.segments and segments > 1.3 * w * h * spp
    # no libm function with differing implementations is on this path on the CPU (uniform environment, atan2f of the
    # same libm on both sides): the per-pixel means agree bit for bit
    assert np.array_equal(img[..., :3].view(np.uint32), ref[..., :3].astype(np.float32).view(np.uint32))
    assert np.all(img[..., 3] == 1.0)


# ---- any scene the host API can describe (tests/scene_file.py -> tests/c/scene_file.h) -----------------------------
def run_scene(harness, tmp_path, variant, scene, rs, w, h, spp):
    from scene_file import write_scene
    sp, out = str(tmp_path / "scene.vrscene"), str(tmp_path / "accum.bin")
    write_scene(sp, scene)
    r = subprocess.run([harness(variant), sp, str(w), str(h), str(spp), str(rs.max_bounces), hex(rs.seed),
                        repr(float(rs.firefly_clamp)), str(int(rs.render_mode)), str(int(rs.pixel_mapping)), out],
                       capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 of " in r.stdout
    segments = int(r.stdout.split(" segments")[0].split()[-1])
    return np.fromfile(out, np.float32).reshape(h, w, 4), segments


def check_scene(oracle, harness, tmp_path, variant, scene, rs, w, h, spp, exact=True):
    rs.total_samples = spp
    img, segments = run_scene(harness, tmp_path, variant, scene, rs, w, h, spp)
    ref, counters = oracle.OracleScene(scene).render(w, h, rs, spp)
    ref = ref.astype(np.float32)
    assert segments == counters.segments
    if exact:
        # both sides evaluate the same f32 operations in the same order and call the same libm here, so even the
        # HDRI lookups (acosf / atan2f) and the 30-degree normal test agree bit for bit on the CPU
        assert np.array_equal(img[..., :3].view(np.uint32), ref[..., :3].view(np.uint32)), \
            float(np.abs(img[..., :3] - ref[..., :3]).max())
    else:
        assert float(np.abs(img[..., :3] - ref[..., :3]).max()) <= 2e-4
    return img


@pytest.mark.parametrize("variant", ["default"])
def test_kernels_on_the_cpu_textured_hdri_dof(oracle, harness, tmp_path, variant):
    from voidray_b200 import scenes
    # configs[0]: albedo texture (bilinear), HDRI environment, thin-lens camera; then with the raw normal map
    scene, st, _ = scenes.config1_mushroom(48, 36, 4)
    img = check_scene(oracle, harness, tmp_path, variant, scene, st.render, 48, 36, 4)
    assert float(img[..., :3].max()) > 0.5
    scene, st, _ = scenes.config1_mushroom(48, 36, 4, normal_map=True)
    check_scene(oracle, harness, tmp_path, variant, scene, st.render, 48, 36, 4)


def test_kernels_on_the_cpu_all_material_kinds(oracle, harness, tmp_path):
    from voidray_b200 import scenes
    # lambertian, metal, dielectric, wood texture + nearest-sampled normal map, two meshes
    scene, st, _ = scenes.config3_materials(48, 27, 4)
    check_scene(oracle, harness, tmp_path, "default", scene, st.render, 48, 27, 4)
    # emission, spheres, ground plane, quads, lambertian_bsdf (the reference's example scenes)
    for fn in (scenes.example_cornell, scenes.example_spheres, scenes.example_material):
        scene, st, _ = fn()
        check_scene(oracle, harness, tmp_path, "default", scene, RenderSettings(total_samples=4, max_bounces=10), 40, 40, 4)


def test_kernels_on_the_cpu_normal_mode_and_reference_pixel_mapping(oracle, harness, tmp_path):
    from voidray_b200 import scenes
    from voidray_b200.scene import PixelMapping, RenderMode
    scene, st, _ = scenes.config1_mushroom(40, 40, 2, dof=False)
    st.render.render_mode = RenderMode.Normal
    check_scene(oracle, harness, tmp_path, "default", scene, st.render, 40, 40, 2)
    st.render.render_mode = RenderMode.Full
    st.render.pixel_mapping = PixelMapping.Reference
    check_scene(oracle, harness, tmp_path, "default", scene, st.render, 40, 40, 2)


@pytest.mark.parametrize("name", ["transparent", "clear", "metallic", "light"])
def test_kernels_on_the_cpu_microfacet(oracle, harness, tmp_path, name):
    # k_shade<false, true>: MicrofacetBSDF through the blanket BSDFMaterial impl (exp / ln / atan / sin / cos: the same
    # libm on both sides here, unlike on the GPU, where tests/test_microfacet.py needs a tolerance)
    from test_microfacet import MATERIALS
    from test_oracle_shading import sphere_scene
    scene = sphere_scene(MATERIALS[name], env=(0.5, 0.5, 0.5))
    rs = RenderSettings(total_samples=8, max_bounces=8)
    w = h = 32
    img, segments = run_scene(harness, tmp_path, "default", scene, rs, w, h, 8)
    ref, counters = oracle.OracleScene(scene).render(w, h, rs, 8)
    ref = ref.astype(np.float32)
    assert segments == counters.segments
    fin = np.isfinite(ref[..., :3]).all(axis=2) & np.isfinite(img[..., :3]).all(axis=2)
    assert fin.mean() > 0.99
    assert np.array_equal(img[..., :3][fin].view(np.uint32), ref[..., :3][fin].view(np.uint32)), \
        float(np.abs(img[..., :3][fin] - ref[..., :3][fin]).max())


def test_fast_division_by_launch_invariants(tmp_path):
    # k_raygen / k_shade divide slots by the pixel count and the tiles per row with a multiply-high (kernels.cuh)
    exe = str(tmp_path / "fastdiv_check")
    csrc = os.path.join(ROOT, "voidray_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=c++20", "-pthread", "-DVR_HOST_SHIM", "-DVR_HOST_SIMT", "-I", os.path.join(ROOT, "tests", "c"),
                        "-I", csrc, os.path.join(ROOT, "tests", "c", "fastdiv_check.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and ", 0 wrong" in r.stdout, r.stdout



def test_rejection_samplers_up_front_equal_the_plain_loops(tmp_path):
    # UnitSphere / UnitCircle / UnitDisc evaluate their first three candidate pairs up front (device_math.cuh
    # Rng::accepted_pair): same values, same draws consumed, same stream afterwards as rand_distr's plain loops
    exe = str(tmp_path / "rng_pairs_check")
    csrc = os.path.join(ROOT, "voidray_b200", "csrc")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DVR_HOST_SHIM", "-I", os.path.join(ROOT, "tests", "c"),
                        "-I", csrc, os.path.join(ROOT, "tests", "c", "rng_pairs_check.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and ", 0 wrong" in r.stdout, r.stdout
