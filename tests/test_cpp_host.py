"""The C++ host mirror (include/voidray.hpp) and its example: compiles and links on CPU; on the GPU the example's
render of the reference's cornell scene is bit-identical to the Python host's (both drive the same library)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")


def _build():
    subprocess.check_call(["make", "-C", EXAMPLES, "-s"])
    return os.path.join(EXAMPLES, "cornell")


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
    buf = np.ascontiguousarray(tgt.read())
    d = 1469598103934665603
    for b in buf.tobytes():
        d = ((d ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert f"{d:016x}" == m.group(1)
    assert tgt.stats().ray_segments == int(m.group(2))
