"""Sanitizer-instrumented mutation fuzzing of the host code that parses untrusted input: the image decoders, the OBJ
loader and the scene flattener (degenerate geometry). The harnesses under tests/fuzz/ are compiled with
-fsanitize=address,undefined against the library's own sources; any memory error or undefined behaviour aborts them.
Short campaigns here (seconds); the same binaries take larger counts for longer runs."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "voidray_b200", "csrc")
FUZZ = os.path.join(ROOT, "tests", "fuzz")
SAN = ["-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
ENV = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1")


def _build(tmp_path, name, sources, extra=()):
    exe = str(tmp_path / name)
    errors = []
    for cxx in dict.fromkeys([os.environ.get("CXX", "g++"), "g++", "/usr/bin/g++", "clang++"]):
        if not shutil.which(cxx):
            continue
        cmd = [cxx, *SAN, "-I", CSRC, *sources, os.path.join(FUZZ, name + ".cpp"), "-o", exe, *extra]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            return exe
        errors.append(r.stderr[-1500:])
        if "asan" not in r.stderr.lower() and "ubsan" not in r.stderr.lower():
            break  # a real compile error, not a missing sanitizer runtime
    else:
        pytest.skip("no compiler here can link the sanitizer runtimes")
    raise AssertionError(errors[-1])


def _corpus(tmp_path):
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    from PIL import Image
    d = tmp_path / "corpus"
    d.mkdir()
    y, x = np.mgrid[0:37, 0:53]
    rgb = np.stack([127 + 100 * np.sin(x / 9.0) * np.cos(y / 11.0), 127 + 90 * np.cos(x / 13.0 + y / 7.0),
                    127 + 110 * np.sin((x + y) / 17.0)], axis=2).astype(np.uint8)
    Image.fromarray(rgb).save(d / "rgb.png")
    Image.fromarray(rgb).quantize(13).save(d / "pal4.png", bits=4)
    Image.fromarray(rgb[:, :, 0]).save(d / "grey.png")
    cv2.imwrite(str(d / "rgb16.png"), rgb.astype(np.uint16) * 257)
    for c in ("raw", "tiff_lzw", "tiff_adobe_deflate", "packbits"):
        Image.fromarray(rgb).save(d / f"{c}.tif", compression=c)
    cv2.imwrite(str(d / "rgb16.tif"), rgb.astype(np.uint16) * 257)
    Image.fromarray(rgb).save(d / "base.jpg", quality=90, subsampling=0)
    Image.fromarray(rgb).save(d / "sub.jpg", quality=90, subsampling=2, restart_marker_blocks=2)
    Image.fromarray(rgb).save(d / "prog.jpg", quality=90, subsampling=2, progressive=True)
    Image.fromarray(rgb).save(d / "rgb.bmp")
    Image.fromarray(rgb).quantize(16).save(d / "pal4.bmp", bits=4)
    Image.fromarray(rgb).save(d / "raw.tga")
    Image.fromarray(rgb).save(d / "rle.tga", compression="tga_rle")
    Image.fromarray(rgb).quantize(40).save(d / "pal.tga", compression="tga_rle")
    Image.fromarray(rgb).save(d / "p6.ppm")
    Image.fromarray((rgb[:, :, 0] > 128).astype(np.uint8) * 255).convert("1").save(d / "p4.pbm")
    (d / "p3.ppm").write_text("P3 4 2 255\n" + " ".join(str(int(v)) for v in rgb[:2, :4].reshape(-1)))
    (d / "t.ff").write_bytes(b"farbfeld" + (5).to_bytes(4, "big") + (3).to_bytes(4, "big") + bytes(range(120)))
    f = (rgb.astype(np.float32) / 16.0) ** 2
    cv2.imwrite(str(d / "t.hdr"), f)
    if hasattr(cv2, "IMWRITE_EXR_COMPRESSION"):
        for name, flag in (("none", 0), ("rle", 1), ("zips", 2), ("zip", 3), ("piz", 4), ("pxr24", 5), ("b44", 6), ("b44a", 7)):
            for half in (0, 1):
                cv2.imwrite(str(d / f"{name}_{half}.exr"), f, [cv2.IMWRITE_EXR_COMPRESSION, flag, cv2.IMWRITE_EXR_TYPE,
                                                               cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT])
    Image.fromarray(rgb).quantize(40).save(d / "a.gif")
    Image.fromarray(rgb).quantize(6).save(d / "b.gif", interlace=True)
    icon = Image.fromarray(np.dstack([rgb[:32, :32], np.full((32, 32), 255, np.uint8)]))
    icon.save(d / "p.ico", sizes=[(32, 32)])
    icon.save(d / "q.ico", sizes=[(16, 16), (32, 32)], bitmap_format="bmp")
    for fmt in ("DXT1", "DXT5"):
        Image.fromarray(np.dstack([rgb, rgb[:, :, 0]])).save(d / f"{fmt}.dds", pixel_format=fmt)
    from test_image_io import _rle8_bmp, _write_tiled_exr, _write_tiled_tiff
    pal_img = Image.fromarray(rgb).quantize(32)
    (d / "rle8.bmp").write_bytes(_rle8_bmp(np.asarray(pal_img), np.asarray(pal_img.getpalette()[:96], np.uint8).reshape(-1, 3), 2))
    _write_tiled_tiff(str(d / "tiled8.tif"), rgb, (32, 16), True, True)
    _write_tiled_tiff(str(d / "tiled16.tif"), rgb.astype(np.uint16) * 257, (16, 16), False, True, big_endian=True)
    for comp in (0, 3):
        _write_tiled_exr(str(d / f"tiled_{comp}.exr"), f[:21, :30], (16, 8), comp, half=bool(comp))
    return str(d)


def test_image_decoders_survive_mutated_files(tmp_path):
    exe = _build(tmp_path, "fuzz_image", [os.path.join(CSRC, "image_io.cpp")], extra=["-fwrapv", "-lz"])
    r = subprocess.run([exe, _corpus(tmp_path), "150", "7"], capture_output=True, text=True, env=ENV, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert "decoded" in r.stdout and "BASE FAIL" not in r.stdout


@pytest.fixture(scope="module")
def scene_build_object(tmp_path_factory):
    """csrc/scene_build.cpp compiled once with the sanitizers, linked into both harnesses that need it."""
    d = tmp_path_factory.mktemp("san")
    obj = str(d / "scene_build.o")
    for cxx in dict.fromkeys([os.environ.get("CXX", "g++"), "g++", "/usr/bin/g++"]):
        if not shutil.which(cxx):
            continue
        # probe: can this compiler link the sanitizer runtimes at all?
        probe = d / "probe.cpp"
        probe.write_text("int main() { return 0; }\n")
        if subprocess.run([cxx, *SAN, str(probe), "-o", str(d / "probe")], capture_output=True).returncode != 0:
            continue
        r = subprocess.run([cxx, *SAN, "-I", CSRC, "-x", "c++", "-c", os.path.join(CSRC, "scene_build.cpp"), "-o", obj],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        return obj
    pytest.skip("no compiler here can link the sanitizer runtimes")


def test_obj_loader_survives_mutated_files(tmp_path, scene_build_object):
    exe = _build(tmp_path, "fuzz_obj", [scene_build_object])
    objs = [os.path.join(ROOT, "assets", n) for n in ("cube.obj", "fancy_monkey.obj")]
    r = subprocess.run([exe, "3", "300", str(tmp_path / "scratch.obj"), *objs], capture_output=True, text=True, env=ENV, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert "loaded" in r.stdout


def test_flatten_survives_degenerate_scenes(tmp_path, scene_build_object):
    exe = _build(tmp_path, "fuzz_flatten", [scene_build_object])
    r = subprocess.run([exe, "5", "600"], capture_output=True, text=True, env=ENV, timeout=600)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert r.stdout.strip().endswith("ok")
