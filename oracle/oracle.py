"""ctypes wrapper of the CPU oracle (oracle/voidray_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. Nothing under voidray_b200/ imports this module. PARITY UNPINNED (see the
header of voidray_oracle.cpp): the reference has no golden vectors for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

from voidray_b200.scene import (GroundPlaneDesc, HDRIEnvironment, MeshData, ObjFile, RenderSettings, Scene,
                                SphereDesc, UniformEnvironment)

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvoidray_oracle.so")

MODE_FAITHFUL, MODE_EARLY_OUT, MODE_BRUTE = 0, 1, 2


class _MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", C.c_float * 3), ("param", C.c_float), ("albedo_tex", C.c_int32),
                ("normal_tex", C.c_int32), ("index", C.c_float), ("roughness", C.c_float), ("metallic", C.c_float),
                ("emittance", C.c_float), ("transparent", C.c_int32)]


class _Settings(C.Structure):
    _fields_ = [("total_samples", C.c_uint32), ("max_bounces", C.c_uint32), ("firefly_clamp", C.c_float),
                ("render_mode", C.c_int32), ("pixel_mapping", C.c_int32), ("traverse_mode", C.c_int32),
                ("seed", C.c_uint64), ("integrator", C.c_int32), ("reserved", C.c_int32)]


class _Counters(C.Structure):
    _fields_ = [("box_tests", C.c_uint64), ("tri_tests", C.c_uint64), ("segments", C.c_uint64)]


@dataclass
class Counters:
    box_tests: int
    tri_tests: int
    segments: int


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "voidray_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.vo_scene_create.restype = C.c_void_p
        lib.vo_schlick.restype = C.c_float
        lib.vo_schlick.argtypes = [C.c_float, C.c_float]
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class OracleScene:
    """Builds the oracle's copy of a voidray_b200.scene.Scene and exposes the reference operations."""

    def __init__(self, scene: Scene):
        lib = load()
        self._lib = lib
        self.scene = scene
        self.handle = C.c_void_p(lib.vo_scene_create())
        h = self.handle
        for tex in scene.textures:
            img = np.ascontiguousarray(tex.image, dtype=F32)
            lib.vo_add_texture(h, _fp(img), C.c_uint32(img.shape[1]), C.c_uint32(img.shape[0]), C.c_int32(int(tex.sample_type)))
        for surf in scene.surfaces:
            if isinstance(surf, ObjFile):
                surf = surf.mesh()      # the oracle gets the Python loader's arrays (assets.load_obj)
            if isinstance(surf, MeshData):
                pos = np.ascontiguousarray(surf.positions, dtype=F32)
                uvs = np.ascontiguousarray(surf.uvs, dtype=F32)
                nrm = np.ascontiguousarray(surf.normals, dtype=F32)
                idx = np.ascontiguousarray(surf.indices, dtype=np.uint32)
                lib.vo_add_mesh(h, _fp(pos), _fp(uvs), _fp(nrm), C.c_uint32(pos.shape[0]), _up(idx), C.c_uint32(idx.size))
            elif isinstance(surf, SphereDesc):
                lib.vo_add_sphere(h, _f3(surf.center), C.c_float(surf.radius))
            elif isinstance(surf, GroundPlaneDesc):
                lib.vo_add_ground_plane(h, C.c_float(surf.height))
            else:
                raise TypeError(surf)
        for m in scene.materials:
            d = _MaterialDesc(int(m.kind), (C.c_float * 3)(*m.color), float(m.param), int(m.albedo_tex), int(m.normal_tex),
                              float(m.index), float(m.roughness), float(m.metallic), float(m.emittance),
                              1 if m.transparent else 0)
            if lib.vo_add_material(h, C.byref(d)) < 0:
                raise ValueError(f"oracle: unknown material kind {m.kind}")
        for o in scene.objects:
            lib.vo_add_object(h, C.c_uint32(o.material), C.c_uint32(o.surface))
        cam = scene.camera
        has_dof = cam.dof is not None
        lib.vo_set_camera(h, _f3(cam.eye), _f3(cam.direction), _f3(cam.up), C.c_float(cam.fov),
                          C.c_int32(1 if has_dof else 0), C.c_float(cam.dof[0] if has_dof else 0.0),
                          _f3(cam.dof[1]) if has_dof else None)
        env = scene.environment
        if isinstance(env, UniformEnvironment):
            lib.vo_set_environment_uniform(h, _f3(env.color))
        elif isinstance(env, HDRIEnvironment):
            img = np.ascontiguousarray(env.image, dtype=F32)
            lib.vo_set_environment_hdri(h, _fp(img), C.c_uint32(img.shape[1]), C.c_uint32(img.shape[0]))
        if lib.vo_commit(h) != 0:
            raise ValueError("oracle: every surface needs an object at its own index (core/scene.rs:183-184)")

    def __del__(self):
        try:
            if self.handle:
                self._lib.vo_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @staticmethod
    def _settings(rs: RenderSettings, traverse_mode: int) -> _Settings:
        return _Settings(int(rs.total_samples), int(rs.max_bounces), float(rs.firefly_clamp), int(rs.render_mode),
                         int(rs.pixel_mapping), int(traverse_mode), int(rs.seed), int(rs.integrator), 0)

    def trace_rays(self, origins, directions, mode: int = MODE_FAITHFUL, details: bool = False):
        o = np.ascontiguousarray(origins, dtype=F32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=F32).reshape(-1, 3)
        n = o.shape[0]
        surface = np.empty(n, np.uint32)
        prim = np.empty(n, np.uint32)
        t = np.empty(n, F32)
        normal = np.empty((n, 3), F32) if details else None
        uv = np.empty((n, 2), F32) if details else None
        front = np.empty(n, np.uint8) if details else None
        c = _Counters()
        self._lib.vo_trace_rays(self.handle, C.c_uint64(n), _fp(o), _fp(d), C.c_int32(mode), _up(surface), _up(prim),
                                _fp(t), _fp(normal), _fp(uv),
                                front.ctypes.data_as(C.POINTER(C.c_uint8)) if details else None, C.byref(c))
        counters = Counters(c.box_tests, c.tri_tests, c.segments)
        if details:
            return surface, prim, t, normal, uv, front, counters
        return surface, prim, t, counters

    def trace_primary(self, width: int, height: int, rs: RenderSettings, sample: int = 0, mode: int = MODE_FAITHFUL):
        n = width * height
        origins = np.empty((n, 3), F32)
        dirs = np.empty((n, 3), F32)
        surface = np.empty(n, np.uint32)
        prim = np.empty(n, np.uint32)
        t = np.empty(n, F32)
        c = _Counters()
        st = self._settings(rs, mode)
        self._lib.vo_trace_primary(self.handle, C.c_uint32(width), C.c_uint32(height), C.byref(st), C.c_uint32(sample),
                                   _fp(origins), _fp(dirs), _up(surface), _up(prim), _fp(t), C.byref(c))
        return origins, dirs, surface, prim, t, Counters(c.box_tests, c.tri_tests, c.segments)

    def sample_radiance(self, width: int, height: int, rs: RenderSettings, pixels, samples, mode: int = MODE_FAITHFUL):
        px = np.ascontiguousarray(pixels, dtype=np.uint32).reshape(-1)
        sm = np.ascontiguousarray(samples, dtype=np.uint32).reshape(-1)
        out = np.empty((px.size, 3), F32)
        st = self._settings(rs, mode)
        self._lib.vo_sample_radiance(self.handle, C.c_uint32(width), C.c_uint32(height), C.byref(st), C.c_uint64(px.size),
                                     _up(px), _up(sm), _fp(out))
        return out

    def render(self, width: int, height: int, rs: RenderSettings, samples: int, accum: Optional[np.ndarray] = None,
               sample_offset: Optional[int] = None, pixel_begin: int = 0, pixel_end: Optional[int] = None,
               n_threads: int = 0, mode: int = MODE_FAITHFUL):
        """iterative_render: adds `samples` camera samples per pixel into accum ((H, W, 4) f32)."""
        if accum is None:
            accum = np.zeros((height, width, 4), F32)
        assert accum.dtype == F32 and accum.flags.c_contiguous and accum.size == width * height * 4
        if pixel_end is None:
            pixel_end = width * height
        if sample_offset is None:
            sample_offset = rs.sample_offset
        if n_threads <= 0:
            n_threads = os.cpu_count() or 1
        c = _Counters()
        st = self._settings(rs, mode)
        self._lib.vo_render(self.handle, C.c_uint32(width), C.c_uint32(height), C.byref(st), C.c_uint32(sample_offset),
                            C.c_uint32(samples), C.c_uint32(pixel_begin), C.c_uint32(pixel_end), _fp(accum),
                            C.c_int32(n_threads), C.byref(c))
        return accum, Counters(c.box_tests, c.tri_tests, c.segments)

    def mesh_tie_rank(self, surface: int) -> np.ndarray:
        n = self.scene.surfaces[surface].n_triangles
        out = np.empty(n, np.uint32)
        self._lib.vo_mesh_tie_rank(self.handle, C.c_uint32(surface), _up(out))
        return out

    def surface_rank(self) -> np.ndarray:
        out = np.empty(len(self.scene.surfaces), np.uint32)
        self._lib.vo_surface_rank(self.handle, _up(out))
        return out

    def global_tie_rank(self, surface: int) -> np.ndarray:
        """Rank in the whole scene's in-order sequence (what the product exports)."""
        order = np.argsort(self.surface_rank(), kind="stable")
        base = 0
        for s in order:
            surf = self.scene.surfaces[s]
            if s == surface:
                return (self.mesh_tie_rank(surface) + np.uint32(base)).astype(np.uint32)
            base += surf.n_triangles if isinstance(surf, (MeshData, ObjFile)) else 1
        raise IndexError(surface)

    def texture_sample(self, texture: int, uv) -> np.ndarray:
        uv = np.ascontiguousarray(uv, dtype=F32).reshape(-1, 2)
        out = np.empty((uv.shape[0], 3), F32)
        self._lib.vo_texture_sample(self.handle, C.c_uint32(texture), C.c_uint64(uv.shape[0]), _fp(uv), _fp(out))
        return out

    def environment_sample(self, directions) -> np.ndarray:
        d = np.ascontiguousarray(directions, dtype=F32).reshape(-1, 3)
        out = np.empty((d.shape[0], 3), F32)
        self._lib.vo_environment_sample(self.handle, C.c_uint64(d.shape[0]), _fp(d), _fp(out))
        return out


def resolve(accum: np.ndarray, scale: float, gamma: float, exposure: float, tonemap: int) -> np.ndarray:
    a = np.ascontiguousarray(accum, dtype=F32)
    h, w = a.shape[0], a.shape[1]
    out = np.empty_like(a)
    load().vo_resolve(_fp(a), C.c_uint32(w), C.c_uint32(h), C.c_float(scale), C.c_float(gamma), C.c_float(exposure),
                      C.c_int32(tonemap), _fp(out))
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.empty(4, np.uint32)
    load().vo_philox4x32_10(_up(c), _up(k), _up(out))
    return out


def rng_draws(seed: int, pixel: int, sample: int, n: int) -> np.ndarray:
    out = np.empty(n, np.uint32)
    load().vo_rng_draws(C.c_uint64(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(n), _up(out))
    return out


def unit_sphere(seed: int, pixel: int, sample: int, n: int) -> np.ndarray:
    out = np.empty((n, 3), F32)
    load().vo_unit_sphere(C.c_uint64(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(n), _fp(out))
    return out


def unit_disc(seed: int, pixel: int, sample: int, n: int) -> np.ndarray:
    out = np.empty((n, 2), F32)
    load().vo_unit_disc(C.c_uint64(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(n), _fp(out))
    return out


def schlick(cosine: float, idx: float) -> float:
    return float(load().vo_schlick(C.c_float(cosine), C.c_float(idx)))
