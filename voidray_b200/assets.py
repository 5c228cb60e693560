"""Asset decoding for the host side: OBJ meshes, image textures, and the closed-form substitutes for
the HDRIs that are missing from the reference checkout (`.MISSING_LARGE_BLOBS`).

In the reference these are third-party crates, not renderer logic:
  obj-rs 0.7.0   load_obj::<TexturedVertex, u32>   (core/mesh.rs:46-74)
  image 0.24.3   image::open(path).to_rgb32f()     (core/texture.rs:36-49, environments.rs:42-55)
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np

F32 = np.float32
ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def asset_path(name: str) -> str:
    return os.path.join(ASSET_DIR, name)


def load_obj(path: str):
    """obj-rs `load_obj::<TexturedVertex, u32>`: every face must be a triangle of v/vt/vn triples;
    each distinct (position, texture, normal) index triple becomes one vertex, in first-seen order;
    `texture` is (u, v, w) and the reference keeps (u, v) (mesh.rs:58)."""
    from .scene import MeshData

    pos, tex, nrm = [], [], []
    triples = {}
    order = []
    indices = []
    with open(path, "r") as f:
        for line in f:
            if not line or line[0] not in "vf":
                continue
            parts = line.split()
            if not parts:
                continue
            tag = parts[0]
            if tag == "v":
                pos.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif tag == "vt":
                tex.append((float(parts[1]), float(parts[2]) if len(parts) > 2 else 0.0))
            elif tag == "vn":
                nrm.append((float(parts[1]), float(parts[2]), float(parts[3])))
            elif tag == "f":
                if len(parts) != 4:
                    raise ValueError(f"{path}: model should be triangulated first to be loaded properly")
                for p in parts[1:]:
                    ids = p.split("/")
                    if len(ids) != 3 or not ids[1] or not ids[2]:
                        raise ValueError(f"{path}: TexturedVertex needs position/texture/normal on every face vertex")
                    key = (int(ids[0]), int(ids[1]), int(ids[2]))
                    # negative indices are relative to the end of the list read so far
                    key = tuple(k - 1 if k > 0 else n + k for k, n in zip(key, (len(pos), len(tex), len(nrm))))
                    idx = triples.get(key)
                    if idx is None:
                        idx = len(order)
                        triples[key] = idx
                        order.append(key)
                    indices.append(idx)
    pos_a = np.asarray(pos, dtype=np.float64).astype(F32)
    tex_a = np.asarray(tex, dtype=np.float64).astype(F32)
    nrm_a = np.asarray(nrm, dtype=np.float64).astype(F32)
    keys = np.asarray(order, dtype=np.int64).reshape(-1, 3)
    return MeshData(
        positions=np.ascontiguousarray(pos_a[keys[:, 0]]),
        uvs=np.ascontiguousarray(tex_a[keys[:, 1]]),
        normals=np.ascontiguousarray(nrm_a[keys[:, 2]]),
        indices=np.asarray(indices, dtype=np.uint32),
    )


def load_obj_native(path: str) -> MeshData:
    """`Mesh::from_file` (core/mesh.rs:46-74) through the library's own OBJ loader (vr_obj_load,
    csrc/scene_build.cpp) — the loader `Scene.add_mesh_from_file` uses at commit. Raises VoidrayError where the
    reference would panic."""
    import ctypes as C

    from . import _lib
    from .scene import MeshData

    lib = _lib.load()
    m = _lib.ObjMeshC()
    _lib.check(lib.vr_obj_load(os.fsencode(path), C.byref(m)))
    try:
        nv, ni = m.n_vertices, m.n_indices
        arr = lambda p, n, t: np.ctypeslib.as_array(p, shape=(n,)).astype(t, copy=True) if n else np.zeros(0, t)  # noqa: E731
        return MeshData(positions=arr(m.positions, 3 * nv, F32).reshape(-1, 3), uvs=arr(m.uvs, 2 * nv, F32).reshape(-1, 2),
                        normals=arr(m.normals, 3 * nv, F32).reshape(-1, 3), indices=arr(m.indices, ni, np.uint32))
    finally:
        lib.vr_obj_free(C.byref(m))


def load_image_native(path: str) -> np.ndarray:
    """`image::open(path).to_rgb32f()` through the library's own decoders (vr_image_load_rgb32f,
    csrc/image_io.cpp): PNG, JPEG, TIFF, Radiance HDR, OpenEXR. Returns (h, w, 3) f32. Raises VoidrayError
    where the reference would panic."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    w, h = C.c_uint32(), C.c_uint32()
    ptr = C.POINTER(C.c_float)()
    _lib.check(lib.vr_image_load_rgb32f(os.fsencode(path), C.byref(w), C.byref(h), C.byref(ptr)))
    try:
        return np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 3)).copy()
    finally:
        lib.vr_image_free(ptr)


def load_image_rgb32f(path: str) -> np.ndarray:
    """`image::open(path).to_rgb32f()`: 8-bit channels / 255, 16-bit / 65535, float passthrough,
    alpha dropped, no sRGB decode. Returns (h, w, 3) f32. Decodes with PIL / OpenCV: the independent
    check of `load_image_native` (tests/test_image_io.py); the benchmark scene recipes feed its arrays to the library."""
    ext = os.path.splitext(path)[1].lower()
    if ext in (".exr", ".hdr", ".pfm"):
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        import cv2

        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(path)
        if img.ndim == 2:
            img = np.repeat(img[:, :, None], 3, axis=2)
        return np.ascontiguousarray(img[:, :, 2::-1].astype(F32))  # BGR(A) -> RGB
    from PIL import Image

    im = Image.open(path)
    if im.mode in ("I;16", "I;16B", "I;16L", "I"):
        a = np.asarray(im).astype(F32) / F32(65535.0)
        return np.ascontiguousarray(np.repeat(a[:, :, None], 3, axis=2))
    if im.mode not in ("RGB", "RGBA"):
        im = im.convert("RGBA" if "A" in im.mode else "RGB")
    a = np.asarray(im)
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a[:, :, :3].astype(F32) / F32(255.0))
    if a.dtype == np.uint16:
        return np.ascontiguousarray(a[:, :, :3].astype(F32) / F32(65535.0))
    return np.ascontiguousarray(a[:, :, :3].astype(F32))


# ---- synthetic lat-long HDRIs ------------------------------------------------------
def _ss(e0, e1, x):
    t = np.clip((x - e0) / (e1 - e0), 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)


def _box(phi, v, p0, p1, v0, v1):
    return _ss(p0, p0 + 0.02, phi) * (1.0 - _ss(p1 - 0.02, p1, phi)) * _ss(v0, v0 + 0.02, v) * (1.0 - _ss(v1 - 0.02, v1, v))


def _mix(a, b, t):
    return a[None, None, :] * (1.0 - t[:, :, None]) + b[None, None, :] * t[:, :, None]


@lru_cache(maxsize=None)
def synth_hdri(name: str, width: int = 2048, height: int = 1024) -> np.ndarray:
    """Closed-form stand-ins for assets/studio.exr and assets/indoor.exr. Row 0 is the zenith
    (environments.rs:85). Only +,-,*,/ and comparisons in f64, so the arrays are bit-identical on
    every machine; peak radiances (20 / 15) exceed firefly_clamp = 3 so the clamp path is exercised."""
    j = (np.arange(height, dtype=np.float64) + 0.5) / height
    i = 2.0 * np.pi * (np.arange(width, dtype=np.float64) + 0.5) / width
    v, phi = np.meshgrid(j, i, indexing="ij")
    A = lambda *c: np.array(c, dtype=np.float64)  # noqa: E731
    if name == "studio":
        img = _mix(0.8 * A(0.90, 0.95, 1.00), 0.3 * A(0.25, 0.22, 0.20), _ss(0.45, 0.55, v))
        img = img + 20.0 * A(1.00, 0.96, 0.90)[None, None, :] * _box(phi, v, 0.6, 1.1, 0.15, 0.35)[:, :, None]
        img = img + 8.0 * A(0.90, 0.95, 1.00)[None, None, :] * _box(phi, v, 3.4, 4.2, 0.20, 0.40)[:, :, None]
        img = img + 4.0 * A(1.0, 1.0, 1.0)[None, None, :] * _box(phi, v, 5.2, 5.5, 0.25, 0.45)[:, :, None]
    elif name == "indoor":
        img = _mix(A(0.20, 0.17, 0.14), A(0.08, 0.06, 0.05), _ss(0.55, 0.65, v))
        img = img + 15.0 * A(1.00, 0.98, 0.95)[None, None, :] * _box(phi, v, 1.2, 2.0, 0.25, 0.50)[:, :, None]
        img = img + 6.0 * A(1.00, 0.85, 0.60)[None, None, :] * (1.0 - _ss(0.04, 0.06, v))[:, :, None]
    else:
        raise KeyError(name)
    out = np.ascontiguousarray(img.astype(F32))
    out.setflags(write=False)
    return out


def transform_mesh(mesh, rotate_y: float = 0.0, translate=(0.0, 0.0, 0.0), scale: float = 1.0):
    """Baked world-space copy (the reference has no instancing or transforms, core/traits.rs:59-63)."""
    from .scene import MeshData

    c, s = np.cos(rotate_y), np.sin(rotate_y)
    R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float64)
    pos = (mesh.positions.astype(np.float64) @ R.T * scale + np.asarray(translate, dtype=np.float64)).astype(F32)
    nrm = (mesh.normals.astype(np.float64) @ R.T).astype(F32)
    return MeshData(np.ascontiguousarray(pos), mesh.uvs, np.ascontiguousarray(nrm), mesh.indices)


def merge_meshes(meshes):
    from .scene import MeshData

    pos, uvs, nrm, idx = [], [], [], []
    base = 0
    for m in meshes:
        pos.append(m.positions)
        uvs.append(m.uvs)
        nrm.append(m.normals)
        idx.append(m.indices.astype(np.uint32) + np.uint32(base))
        base += m.positions.shape[0]
    return MeshData(np.concatenate(pos), np.concatenate(uvs), np.concatenate(nrm), np.concatenate(idx))
