"""Host-side mirror of the reference's render driver, over the C ABI (include/voidray_cuda.h).

  SceneAcceleration        core/scene.rs:60-70,163-185  (device-resident flattened scene)
  RenderTarget             render/target.rs:80-299      (the RGBA f32 accumulation buffer, on the device)
  iterative_render         render/iterative.rs:11-55
  PostProcessingData/Pass  render/post_process.rs:19-86
  Renderer / RenderAction  render/renderer.rs:165-258, RenderThread::one_shot :35-125
"""
from __future__ import annotations

import ctypes as C
import enum
import threading
import time
import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import MaterialDescC, RenderSettingsC, StatsC, check, fptr, uptr
from .scene import (GroundPlaneDesc, HDRIEnvironment, MeshData, ObjFile, RenderSettings, Scene, Settings,
                    SphereDesc, UniformEnvironment)

F32 = np.float32


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Context:
    """One CUDA device + stream (vr_context). `stream` is a raw cudaStream_t handle (int), e.g.
    `torch.cuda.current_stream().cuda_stream`, or None for a library-owned stream."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = _lib.load()
        self.handle = C.c_void_p()
        check(self._lib.vr_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self.handle)))
        self.device = int(device)
        self.devices = [int(device)]
        self._children = weakref.WeakSet()  # live SceneAccelerations: closed before the context is (any close order)

    @classmethod
    def multi(cls, devices) -> "Context":
        """A device group in this process (vr_context_create_multi): scenes are replicated to every device at commit,
        accumulate shards its sample range over them, read / resolve sum the shards on devices[0] over peer memory."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self.handle = C.c_void_p()
        ids = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        check(self._lib.vr_context_create_multi(ids, len(devices), C.byref(self.handle)))
        self.device = int(devices[0])
        self.devices = [int(d) for d in devices]
        self._children = weakref.WeakSet()
        return self

    def close(self):
        if self.handle:
            for child in list(self._children):
                child.close()
            self._lib.vr_context_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class SceneAcceleration:
    """The committed, device-resident scene (Accelerable::build_acceleration, scene.rs:163-179)."""

    def __init__(self, scene: Scene, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._lib = _lib.load()
        lib = self._lib
        self.handle = C.c_void_p()
        check(lib.vr_scene_create(self.ctx.handle, C.byref(self.handle)))
        self._children = weakref.WeakSet()  # live RenderTargets: ended before the scene is destroyed
        self.ctx._children.add(self)
        self.scene = scene
        out = C.c_uint32()
        for tex in scene.textures:
            img = np.ascontiguousarray(tex.image, dtype=F32)
            check(lib.vr_scene_add_texture_rgb32f(self.handle, fptr(img), img.shape[1], img.shape[0],
                                                  int(tex.sample_type), C.byref(out)))
        for surf in scene.surfaces:
            if isinstance(surf, MeshData):
                pos = np.ascontiguousarray(surf.positions, dtype=F32)
                uvs = np.ascontiguousarray(surf.uvs, dtype=F32)
                nrm = np.ascontiguousarray(surf.normals, dtype=F32)
                idx = np.ascontiguousarray(surf.indices, dtype=np.uint32)
                check(lib.vr_scene_add_mesh(self.handle, fptr(pos), fptr(uvs), fptr(nrm), pos.shape[0], uptr(idx),
                                            idx.size, C.byref(out)))
            elif isinstance(surf, ObjFile):
                check(lib.vr_scene_add_mesh_from_obj_file(self.handle, surf.path.encode(), C.byref(out), None, None))
            elif isinstance(surf, SphereDesc):
                check(lib.vr_scene_add_sphere(self.handle, _f3(surf.center), float(surf.radius), C.byref(out)))
            elif isinstance(surf, GroundPlaneDesc):
                check(lib.vr_scene_add_ground_plane(self.handle, float(surf.height), C.byref(out)))
            else:
                raise TypeError(f"unknown surface {surf!r}")
        for m in scene.materials:
            desc = MaterialDescC(int(m.kind), (C.c_float * 3)(*m.color), float(m.param), int(m.albedo_tex),
                                 int(m.normal_tex), float(m.index), float(m.roughness), float(m.metallic),
                                 float(m.emittance), 1 if m.transparent else 0)
            check(lib.vr_scene_add_material(self.handle, C.byref(desc), C.byref(out)))
        for o in scene.objects:
            check(lib.vr_scene_add_object(self.handle, o.material, o.surface, C.byref(out)))
        cam = scene.camera
        has_dof = cam.dof is not None
        check(lib.vr_scene_set_camera(self.handle, _f3(cam.eye), _f3(cam.direction), _f3(cam.up), float(cam.fov),
                                      1 if has_dof else 0, float(cam.dof[0]) if has_dof else 0.0,
                                      _f3(cam.dof[1]) if has_dof else None))
        env = scene.environment
        if isinstance(env, UniformEnvironment):
            check(lib.vr_scene_set_environment_uniform(self.handle, _f3(env.color)))
        elif isinstance(env, HDRIEnvironment):
            img = np.ascontiguousarray(env.image, dtype=F32)
            check(lib.vr_scene_set_environment_hdri_rgb32f(self.handle, fptr(img), img.shape[1], img.shape[0]))
        t0 = time.perf_counter()
        check(lib.vr_scene_commit(self.handle))
        self.commit_seconds = time.perf_counter() - t0

    def commit(self):
        """Re-run vr_scene_commit (flatten + BVH build + upload) on the already-described scene."""
        t0 = time.perf_counter()
        check(self._lib.vr_scene_commit(self.handle))
        self.commit_seconds = time.perf_counter() - t0

    def info(self) -> dict:
        i = _lib.SceneInfoC()
        check(self._lib.vr_scene_get_info(self.handle, C.byref(i)))
        return {k: getattr(i, k) for k, _ in _lib.SceneInfoC._fields_ if k != "reserved"}

    # ---- gates ----
    def trace_rays(self, origins: np.ndarray, directions: np.ndarray):
        o = np.ascontiguousarray(origins, dtype=F32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=F32).reshape(-1, 3)
        n = o.shape[0]
        surface = np.empty(n, np.uint32)
        prim = np.empty(n, np.uint32)
        t = np.empty(n, F32)
        check(self._lib.vr_debug_trace_rays(self.handle, n, fptr(o), fptr(d), uptr(surface), uptr(prim), fptr(t)))
        return surface, prim, t

    def tie_ranks(self, surface: int) -> np.ndarray:
        n = self.scene.surfaces[surface].n_triangles
        out = np.empty(n, np.uint32)
        check(self._lib.vr_debug_tie_ranks(self.handle, int(surface), uptr(out), n))
        return out

    def texture_sample(self, texture: int, uv: np.ndarray) -> np.ndarray:
        uv = np.ascontiguousarray(uv, dtype=F32).reshape(-1, 2)
        out = np.empty((uv.shape[0], 3), F32)
        check(self._lib.vr_debug_texture_sample(self.handle, int(texture), uv.shape[0], fptr(uv), fptr(out)))
        return out

    def environment_sample(self, directions: np.ndarray) -> np.ndarray:
        d = np.ascontiguousarray(directions, dtype=F32).reshape(-1, 3)
        out = np.empty((d.shape[0], 3), F32)
        check(self._lib.vr_debug_environment_sample(self.handle, d.shape[0], fptr(d), fptr(out)))
        return out

    def close(self):
        if self.handle:
            for child in list(self._children):
                child.close()
            self._lib.vr_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class RendererStats:
    samples_done: int
    total_samples: int
    camera_samples: int
    ray_segments: int
    seconds: float
    device_ms: float
    trace_ms: float
    trace_launches: int
    kernel_launches: int


class RenderTarget:
    """The accumulation target (CpuRenderTarget, render/target.rs:80-131): a zeroed W x H RGBA f32
    buffer holding partial sums already divided by total_samples; lives in HBM."""

    def __init__(self, scene: SceneAcceleration, dimensions: Tuple[int, int], settings: RenderSettings):
        self._lib = _lib.load()
        self.scene = scene
        self.dimensions = (int(dimensions[0]), int(dimensions[1]))
        self.settings = settings
        cs = RenderSettingsC(int(settings.total_samples), int(settings.max_bounces), float(settings.firefly_clamp),
                             int(settings.render_mode), int(settings.pixel_mapping), int(settings.integrator),
                             int(settings.seed),
                             int(settings.sample_offset), int(settings.max_paths_in_flight))
        self.handle = C.c_void_p()
        check(self._lib.vr_render_begin(scene.handle, self.dimensions[0], self.dimensions[1], C.byref(cs),
                                        C.byref(self.handle)))
        scene._children.add(self)

    @property
    def n_pixels(self) -> int:
        return self.dimensions[0] * self.dimensions[1]

    def clear(self):  # target.rs:284-290
        check(self._lib.vr_render_clear(self.handle))

    def accumulate(self, samples: int):
        check(self._lib.vr_render_accumulate(self.handle, int(samples)))

    def cancel(self):
        check(self._lib.vr_render_cancel(self.handle))

    def stats(self) -> RendererStats:
        s = StatsC()
        check(self._lib.vr_render_stats(self.handle, C.byref(s)))
        return RendererStats(s.samples_done, s.total_samples, s.camera_samples, s.ray_segments, s.seconds,
                             s.device_ms, s.trace_ms, s.trace_launches, s.kernel_launches)

    def read(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Copy the accumulation buffer to the host: (H, W, 4) f32."""
        if out is None:
            out = np.empty((self.dimensions[1], self.dimensions[0], 4), F32)
        check(self._lib.vr_render_read_accum(self.handle, fptr(out)))
        return out

    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(self._lib.vr_render_accum_device_ptr(self.handle, C.byref(p)))
        return int(p.value)

    def as_torch(self):
        """Zero-copy torch view of the accumulation buffer (for torch.distributed / NCCL reduces)."""
        import torch

        class _Iface:
            pass

        holder = _Iface()
        holder.__cuda_array_interface__ = {"shape": (self.n_pixels * 4,), "typestr": "<f4",
                                           "data": (self.device_ptr(), False), "version": 3, "strides": None}
        t = torch.as_tensor(holder, device=f"cuda:{self.scene.ctx.device}")
        t._voidray_owner = self  # keep the allocation alive
        return t

    def resolve(self, scale: float, gamma: float, exposure: float, tonemap: int,
                out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.dimensions[1], self.dimensions[0], 4), F32)
        check(self._lib.vr_render_resolve(self.handle, float(scale), float(gamma), float(exposure), int(tonemap),
                                          fptr(out)))
        return out

    # ---- multi-GPU: peer-memory reduce (include/voidray_cuda.h) ----
    IPC_HANDLE_BYTES = 64

    def export_accum_handle(self) -> bytes:
        """CUDA IPC handle of the accumulation buffer, to be sent to the root rank."""
        buf = (C.c_uint8 * self.IPC_HANDLE_BYTES)()
        check(self._lib.vr_render_export_accum(self.handle, buf))
        return bytes(buf)

    @staticmethod
    def _pack_handles(handles):
        blob = b"".join(handles)
        return (C.c_uint8 * len(blob)).from_buffer_copy(blob) if blob else None

    def reduce_peers(self, handles) -> None:
        """accum += the peers' accumulation buffers (other processes' GPUs, read over NVLink), in list order."""
        check(self._lib.vr_render_reduce_peers(self.handle, self._pack_handles(handles), len(handles)))

    def resolve_peers(self, handles, scale: float, gamma: float, exposure: float, tonemap: int,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """Fused reduce + tonemap: one kernel sums own + peers and resolves; accum is left untouched."""
        if out is None:
            out = np.empty((self.dimensions[1], self.dimensions[0], 4), F32)
        check(self._lib.vr_render_resolve_peers(self.handle, self._pack_handles(handles), len(handles), float(scale),
                                                float(gamma), float(exposure), int(tonemap), fptr(out)))
        return out

    def reduce_peer_targets(self, peers) -> None:
        """The same for RenderTargets living in this process."""
        ptrs = (C.c_void_p * max(1, len(peers)))(*[p.device_ptr() for p in peers])
        check(self._lib.vr_render_reduce_peer_ptrs(self.handle, ptrs, len(peers)))

    def resolve_peer_targets(self, peers, scale: float, gamma: float, exposure: float, tonemap: int,
                             out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.dimensions[1], self.dimensions[0], 4), F32)
        ptrs = (C.c_void_p * max(1, len(peers)))(*[p.device_ptr() for p in peers])
        check(self._lib.vr_render_resolve_peer_ptrs(self.handle, ptrs, len(peers), float(scale), float(gamma),
                                                    float(exposure), int(tonemap), fptr(out)))
        return out

    # ---- gates ----
    def trace_primary(self, sample: int = 0):
        n = self.n_pixels
        surface = np.empty(n, np.uint32)
        prim = np.empty(n, np.uint32)
        t = np.empty(n, F32)
        check(self._lib.vr_debug_trace_primary(self.handle, int(sample), uptr(surface), uptr(prim), fptr(t)))
        return surface, prim, t

    def sample_radiance(self, pixels: np.ndarray, samples: np.ndarray) -> np.ndarray:
        px = np.ascontiguousarray(pixels, dtype=np.uint32).reshape(-1)
        sm = np.ascontiguousarray(samples, dtype=np.uint32).reshape(-1)
        assert px.size == sm.size
        out = np.empty((px.size, 3), F32)
        check(self._lib.vr_debug_sample_radiance(self.handle, px.size, uptr(px), uptr(sm), fptr(out)))
        return out

    def close(self):
        if self.handle:
            self._lib.vr_render_end(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def iterative_render(target: RenderTarget, scene: SceneAcceleration, settings: RenderSettings, samples: int) -> None:
    """render/iterative.rs:11-55. `scene` and `settings` are bound to the target at creation (the
    reference passes them on every call; they must be the same objects here)."""
    assert target.scene is scene
    target.accumulate(samples)


@dataclass
class PostProcessingData:  # render/post_process.rs:19-25, shaders/post_process.glsl:10-15
    scale: float
    gamma: float
    exposure: float
    tonemap: int


class PostProcessingPass:
    """render/post_process.rs:27-86 — here over the whole target, not the fixed 1024x1024 dispatch."""

    def render(self, src: RenderTarget, data: PostProcessingData, dst: Optional[np.ndarray] = None) -> np.ndarray:
        return src.resolve(data.scale, data.gamma, data.exposure, data.tonemap, dst)


class RenderAction(enum.Enum):  # render/renderer.rs:165-174
    Render = 0
    Continuous = 1
    Rebuild = 2
    Cancel = 3


class Renderer:
    """render/renderer.rs:155-258 with RenderThread::one_shot (:35-125): a render thread clears the
    target, times one sample per pixel, picks samples_per_frame = update_frequency / t_1spp and loops
    until total_samples, polling for Cancel between batches."""

    def __init__(self, scene: Scene, settings: Settings, dimensions: Tuple[int, int], ctx: Optional[Context] = None):
        self.scene = scene
        self.settings = settings
        self.dimensions = dimensions
        self.ctx = ctx
        self.target: Optional[RenderTarget] = None
        self._thread: Optional[threading.Thread] = None
        self._lock = threading.Lock()
        self._currently_rendering = False
        self._samples = (0, 0)
        self._time: Optional[Tuple[float, Optional[float]]] = None
        self._remaining: Optional[float] = None
        self._cancel = threading.Event()
        self._error: Optional[BaseException] = None

    def _one_shot(self):
        try:
            with self._lock:
                self._currently_rendering = True
                start = time.perf_counter()
                self._time = (start, None)
            accel = self.scene.build_acceleration(self.ctx)  # renderer.rs:58
            rs = self.settings.render
            self.target = RenderTarget(accel, self.dimensions, rs)  # clear, renderer.rs:55
            samples, total = 0, rs.total_samples
            t0 = time.perf_counter()
            iterative_render(self.target, accel, rs, 1)  # renderer.rs:67
            samples += 1
            single = time.perf_counter() - t0
            with self._lock:
                self._samples = (samples, total)
                self._remaining = single * (total - 1)
            spf = int(rs.update_frequency / single) if single > 0 else total
            spf = min(max(spf, 1), max(total - samples, 1))  # renderer.rs:78-81
            while samples < total:
                # renderer.rs:86-91 passes samples_per_frame (overshooting on the last batch); the sum
                # stays normalised by total_samples only if exactly `total` samples are drawn, so this
                # mirror draws delta_samples.
                delta = min(spf, total - samples)
                iterative_render(self.target, accel, rs, delta)
                samples += delta
                elapsed = time.perf_counter() - start
                with self._lock:
                    self._samples = (samples, total)
                    self._remaining = elapsed / samples * (total - samples)
                if self._cancel.is_set():  # renderer.rs:101-106
                    break
        except BaseException as e:  # surfaced by join()
            self._error = e
            if isinstance(e, _lib.RenderCancelled) and self.target is not None:
                # a cancelled accumulate keeps the whole batches it finished: the display scale total / done
                # (post_process) must see them
                with self._lock:
                    self._samples = (int(self.target.stats().samples_done), self._samples[1] or self.settings.render.total_samples)
        finally:
            with self._lock:
                self._currently_rendering = False
                if self._time is not None:
                    self._time = (self._time[0], time.perf_counter())
                self._remaining = None

    def execute(self, action: RenderAction):
        if action == RenderAction.Render:
            if self.currently_rendering():
                raise RuntimeError(f"invalid action {action}")  # renderer.rs:203-205 panics
            self._cancel.clear()
            self._error = None
            self._thread = threading.Thread(target=self._one_shot, daemon=True)
            self._thread.start()
        elif action == RenderAction.Cancel:
            if self._thread is None:
                raise RuntimeError(f"invalid action {action}")  # renderer.rs:229-231
            self._cancel.set()
            if self.target is not None:
                self.target.cancel()
        else:
            raise RuntimeError(f"invalid action {action}")  # Continuous/Rebuild are stubs in the reference

    def join(self):
        if self._thread is not None:
            self._thread.join()
        if self._error is not None and not isinstance(self._error, _lib.RenderCancelled):
            raise self._error

    def currently_rendering(self) -> bool:
        with self._lock:
            return self._currently_rendering

    def samples(self) -> Tuple[int, int]:
        with self._lock:
            return self._samples

    def elapsed_time(self) -> float:
        with self._lock:
            if self._time is None:
                return 0.0
            start, end = self._time
            return (end if end is not None else time.perf_counter()) - start

    def remaining_time(self) -> Optional[float]:
        with self._lock:
            return self._remaining

    def post_process(self) -> np.ndarray:
        """voidray_app/src/main.rs:68-90: scale = total/done (0 if not normal), then the tonemap pass."""
        done, total = self.samples()
        scale = float(total) / float(done) if done > 0 else 0.0
        cm = self.settings.color_management
        return PostProcessingPass().render(self.target, PostProcessingData(scale, cm.gamma, cm.exposure, int(cm.tonemap)))
