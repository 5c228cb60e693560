"""The CUDA kernels themselves on the CPU: tests/c/wavefront_host.cpp compiles voidray_b200/csrc/kernels.cu (k_raygen,
k_trace, k_shade, k_accumulate as nvcc sees them) through tests/c/host_shim.h — one OS thread per CUDA thread, warp
intrinsics over a per-warp barrier, real atomics — and runs one wavefront batch of a small frame. Checked here:
every wavefront hit equals the single-ray traversal (inside the harness), and the accumulated image equals the oracle's
render of the same scene with the same seed. This is the radiance gate of tests/test_gpu_radiance.py without a GPU, and
the only execution the experiment variants' warp-level code (-DVR_TRACE_CHUNK claims, -DVR_BVH4 step) gets before one."""
import os
import subprocess

import numpy as np
import pytest

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Camera, Environments, Materials, RenderSettings, Scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "default": [],
    "refill24": ["-DVR_REFILL_THRESHOLD=24"],
    "chunk": ["-DVR_TRACE_CHUNK", "-DVR_LEAF_COMPACT"],
    "chunk_r24": ["-DVR_TRACE_CHUNK", "-DVR_LEAF_COMPACT", "-DVR_REFILL_THRESHOLD=24"],
    "chunk_r32": ["-DVR_TRACE_CHUNK", "-DVR_LEAF_COMPACT", "-DVR_REFILL_THRESHOLD=32"],
    "bvh4": ["-DVR_BVH4", "-DVR_NODE_STEPS=2"],
    "bvh4_nosort_chunk": ["-DVR_BVH4", "-DVR_BVH4_NOSORT", "-DVR_NODE_STEPS=1", "-DVR_TRACE_CHUNK", "-DVR_LEAF_COMPACT",
                          "-DVR_REFILL_THRESHOLD=20"],
    "stack8_tri48": ["-DVR_SMEM_STACK=8", "-DVR_TRI48"],
}
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


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_kernels_on_the_cpu_match_the_oracle(oracle, harness, tmp_path, variant):
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
