"""include/voidray_cuda.h is a C header: a strict C99 translation unit includes it, links the library and calls the
host-only entry points; the struct sizes it sees equal the ctypes mirror's (voidray_b200/_lib.py)."""
import ctypes as C
import os
import re
import subprocess

from voidray_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_c99_and_host_calls_work(tmp_path):
    exe = str(tmp_path / "abi_host")
    libdir = os.path.join(ROOT, "voidray_b200")
    _lib.load()  # the library is built
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_host.c"), "-L", libdir, "-lvoidray_cuda", f"-Wl,-rpath,{libdir}",
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "assets", "uv_test.png")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "256 x 256 first texel" in r.stdout
    m = re.search(r"sizeof\(settings\) (\d+) sizeof\(stats\) (\d+) sizeof\(material\) (\d+)", r.stdout)
    assert m, r.stdout
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (C.sizeof(_lib.RenderSettingsC), C.sizeof(_lib.StatsC),
                                                                   C.sizeof(_lib.MaterialDescC))


def test_unorm8_conversion_of_the_tex8_experiment_is_exact(tmp_path):
    exe = str(tmp_path / "unorm8_exact")
    r = subprocess.run(["gcc", "-O1", "-ffp-contract=off", os.path.join(ROOT, "tests", "c", "unorm8_exact.c"), "-lm", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "0 of 256 values differ" in r.stdout, r.stdout
