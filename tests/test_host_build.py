"""Host half of vr_scene_commit that needs no GPU: the reference-order leaf sequence behind the tie ranks
(csrc/scene_build.cpp `reference_leaf_order`, core/bvh.rs:48-130) against a direct numpy restatement, on the
serial path and on the multi-threaded top-of-tree path (forced by VOIDRAY_PAR_MIN)."""
import ctypes as C
import os

import numpy as np
import pytest

from voidray_b200 import _lib

F32 = np.float32


def _total_key(x: np.ndarray) -> np.ndarray:
    i = x.view(np.int32).astype(np.int64)
    return np.where(i < 0, i ^ 0x7FFFFFFF, i)  # f32::total_cmp order


def _naive_order(cen: np.ndarray) -> np.ndarray:
    order = np.arange(len(cen))
    stack = [(0, len(cen))]
    while stack:
        lo, hi = stack.pop()
        if hi - lo < 2:
            continue
        c = cen[order[lo:hi]]
        spread = c.max(0) - c.min(0)
        if spread[0] > spread[1] and spread[0] > spread[2]:
            axis = 0
        elif spread[1] > spread[0] and spread[1] > spread[2]:
            axis = 1
        else:
            axis = 2
        order[lo:hi] = order[lo:hi][np.argsort(_total_key(c[:, axis]), kind="stable")]
        mid = lo + (hi - lo) // 2
        stack += [(lo, mid), (mid, hi)]
    return order


def _native_order(boxes: np.ndarray, par_min=None) -> np.ndarray:
    lib = _lib.load()
    boxes = np.ascontiguousarray(boxes, F32)
    out = np.empty(len(boxes), np.uint32)
    old = os.environ.get("VOIDRAY_PAR_MIN")
    if par_min is not None:
        os.environ["VOIDRAY_PAR_MIN"] = str(par_min)
    try:
        _lib.check(lib.vr_debug_reference_leaf_order(_lib.fptr(boxes), len(boxes), out.ctypes.data_as(C.POINTER(C.c_uint32))))
    finally:
        if par_min is not None:
            if old is None:
                del os.environ["VOIDRAY_PAR_MIN"]
            else:
                os.environ["VOIDRAY_PAR_MIN"] = old
    return out


def _boxes(n, mode, seed):
    rng = np.random.default_rng(seed)
    if mode == "random":
        c = rng.uniform(-5, 5, (n, 3)).astype(F32)
        h = rng.uniform(0, 0.3, (n, 3)).astype(F32)
    elif mode == "ties":  # a handful of distinct coordinates, signed zeros: stability decides almost everything
        c = (rng.integers(0, 7, (n, 3)) - 3).astype(F32)
        c[rng.random((n, 3)) < 0.02] = F32(-0.0)
        h = np.full((n, 3), 0.5, F32)
    else:  # lattice
        c = (rng.integers(0, 1000, (n, 3)) * 0.25 - 100.0).astype(F32)
        h = rng.uniform(0, 0.3, (n, 3)).astype(F32)
    return np.concatenate([c - h, c + h], axis=1).astype(F32)


@pytest.mark.parametrize("mode", ["random", "ties", "lattice"])
@pytest.mark.parametrize("n", [1, 2, 3, 25, 1000, 40000])
def test_leaf_order_matches_direct_restatement(n, mode):
    boxes = _boxes(n, mode, n)
    cen = ((boxes[:, :3] + boxes[:, 3:]) / F32(2.0)).astype(F32)
    want = _naive_order(cen)
    assert np.array_equal(_native_order(boxes), want)
    assert np.array_equal(_native_order(boxes, par_min=64), want)


@pytest.mark.parametrize("mode", ["random", "ties"])
def test_leaf_order_parallel_top_equals_serial(mode):
    boxes = _boxes(400000, mode, 11)
    serial = _native_order(boxes, par_min=10**9)
    assert np.array_equal(_native_order(boxes, par_min=1 << 16), serial)
    assert sorted(serial.tolist()) == list(range(400000))


# ---- the native OBJ loader (vr_obj_load = the loader behind vr_scene_add_mesh_from_obj_file) ----------------------
from voidray_b200 import assets  # noqa: E402

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


@pytest.mark.parametrize("name", ["cube.obj", "mushroom.obj", "fancy_monkey.obj", "mossy_ground.obj", "material_testing_stand.obj"])
def test_native_obj_loader_matches_python_loader(name):
    # obj-rs 0.7.0 `load_obj::<TexturedVertex, u32>`: one vertex per distinct v/vt/vn triple in first-seen order
    a = assets.load_obj_native(os.path.join(ASSETS, name))
    b = assets.load_obj(os.path.join(ASSETS, name))
    assert np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.positions, b.positions) and np.array_equal(a.uvs, b.uvs) and np.array_equal(a.normals, b.normals)


def test_native_obj_loader_edge_cases(tmp_path):
    def load(text):
        p = tmp_path / "t.obj"
        p.write_text(text)
        return assets.load_obj_native(str(p))

    base = "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\n"
    m = load(base + "f 1/1/1 2/2/1 3/3/1\n")
    assert m.indices.tolist() == [0, 1, 2] and m.positions.shape == (3, 3)
    # negative (relative) indices, a very long comment line, CRLF line ends, other statements ignored
    m2 = load("# " + "x" * 5000 + "\r\nmtllib a.mtl\r\no thing\r\n" + base.replace("\n", "\r\n") +
              "usemtl m\r\ns off\r\nf -3/-3/-1 -2/-2/-1 -1/-1/-1\r\n")
    assert np.array_equal(m2.positions, m.positions) and np.array_equal(m2.uvs, m.uvs) and m2.indices.tolist() == [0, 1, 2]
    # shared corners are deduplicated, distinct triples are not
    m3 = load(base + "f 1/1/1 2/2/1 3/3/1\nf 1/1/1 3/3/1 2/1/1\n")
    assert m3.indices.tolist() == [0, 1, 2, 0, 2, 3] and m3.positions.shape == (4, 3)
    for bad, why in [(base + "f 1/1/1 2/2/1 3/3/1 1/2/1\n", "triangulated"), (base + "f 1 2 3\n", "position/texture/normal"),
                     (base + "f 1/1/1 2/2/1 9/3/1\n", "missing position"), ("f 1/1/1 1/1/1 1/1/1\n", "missing position"),
                     (base + "f 1/1/1 2/2/1\n", "triangulated"), (base + "f 1/7/1 2/2/1 3/3/1\n", "missing texture"),
                     (base + "v 1 2\n", "malformed v")]:
        with pytest.raises(_lib.VoidrayError, match=why):
            load(bad)
    with pytest.raises(_lib.VoidrayError, match="cannot open"):
        assets.load_obj_native(str(tmp_path / "missing.obj"))


def _library_digest(name):
    m = assets.load_obj_native(os.path.join(ASSETS, name))
    lib = _lib.load()
    n = len(m.positions)
    pos = np.ascontiguousarray(m.positions, F32)
    uv = np.zeros((n, 2), F32)  # filler attributes (the digest covers nodes and intersection records)
    nrm = np.tile(np.array([0, 1, 0], F32), (n, 1))
    idx = np.ascontiguousarray(m.indices, np.uint32).ravel()
    digest, nodes, depth, ms = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0), C.c_double(0)
    _lib.check(lib.vr_debug_flatten_mesh_digest(_lib.fptr(pos), _lib.fptr(uv), _lib.fptr(nrm), n,
                                                idx.ctypes.data_as(C.POINTER(C.c_uint32)), idx.size, C.byref(digest),
                                                C.byref(nodes), C.byref(depth), C.byref(ms)))
    return f"{digest.value:016x}", nodes.value, depth.value


@pytest.mark.parametrize("name,want,nodes", [("fancy_monkey.obj", "cbf0e27fea44afdd", None), ("mushroom.obj", "88b409cb71cbab75", 2530),
                                             ("mossy_ground.obj", "8301c250a458b25e", 8382),
                                             ("material_testing_stand.obj", "66408269f0926fcd", 19306)])
def test_shipped_builder_digest(name, want, nodes):
    # The flatten inside libvoidray_cuda.so itself (host-only gate): the bytes the device receives are those of the
    # builder the GPU numbers in profiles/ were measured with, whatever the thread timing (three runs).
    runs = [_library_digest(name) for _ in range(3)]
    assert runs[0] == runs[1] == runs[2]
    assert runs[0][0] == want
    assert nodes is None or runs[0][1] == nodes
    assert runs[0][2] <= 32  # the traversal stack depth
