"""The C-ABI library loads and exports every symbol include/voidray_cuda.h declares; without a GPU the
product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "voidray_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("vr_scene_commit", "vr_render_accumulate", "vr_render_resolve", "vr_render_cancel",
                 "vr_scene_add_mesh", "vr_scene_add_material", "vr_debug_trace_primary"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol():
    from voidray_b200 import _lib
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"libvoidray_cuda.so does not export {name}"
    assert lib.vr_abi_version() == 2
    # the ctypes table binds exactly the declared functions
    bound = set(_lib.SIGNATURES) | set(_lib.NON_STATUS)
    assert bound == set(declared_functions())


def test_struct_layouts_match_the_header():
    from voidray_b200 import _lib
    assert C.sizeof(_lib.MaterialDescC) == 48
    assert C.sizeof(_lib.RenderSettingsC) == 40 and _lib.RenderSettingsC.seed.offset == 24
    assert C.sizeof(_lib.StatsC) == 64


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_gpu_is_a_loud_error_not_a_fallback():
    from voidray_b200 import _lib
    from voidray_b200.render import Context
    with pytest.raises(_lib.VoidrayError) as e:
        Context(0)
    assert e.value.status == _lib.VR_ERR_CUDA
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "voidray_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
