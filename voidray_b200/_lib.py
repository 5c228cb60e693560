"""ctypes binding of libvoidray_cuda.so (include/voidray_cuda.h). There is no fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VOIDRAY_CUDA_LIB") or os.path.join(_HERE, "libvoidray_cuda.so")

VR_OK, VR_ERR_INVALID, VR_ERR_CUDA, VR_ERR_OOM, VR_ERR_CANCELLED = 0, -1, -2, -3, -4


class VoidrayError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"voidray_cuda status {status}: {message}")
        self.status = status


class RenderCancelled(VoidrayError):
    pass


class MaterialDescC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", C.c_float * 3), ("param", C.c_float), ("albedo_tex", C.c_int32),
                ("normal_tex", C.c_int32), ("index", C.c_float), ("roughness", C.c_float), ("metallic", C.c_float),
                ("emittance", C.c_float), ("transparent", C.c_int32)]


class ObjMeshC(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("n_indices", C.c_uint32), ("positions", C.POINTER(C.c_float)),
                ("uvs", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)), ("indices", C.POINTER(C.c_uint32))]


class RenderSettingsC(C.Structure):
    _fields_ = [("total_samples", C.c_uint32), ("max_bounces", C.c_uint32), ("firefly_clamp", C.c_float),
                ("render_mode", C.c_int32), ("pixel_mapping", C.c_int32), ("integrator", C.c_int32),
                ("seed", C.c_uint64), ("sample_offset", C.c_uint32), ("max_paths_in_flight", C.c_uint32)]


class StatsC(C.Structure):
    _fields_ = [("samples_done", C.c_uint32), ("total_samples", C.c_uint32), ("camera_samples", C.c_uint64),
                ("ray_segments", C.c_uint64), ("seconds", C.c_double), ("device_ms", C.c_double),
                ("trace_ms", C.c_double), ("trace_launches", C.c_uint64), ("kernel_launches", C.c_uint64)]


class SceneInfoC(C.Structure):
    _fields_ = [("n_triangles", C.c_uint32), ("n_bvh_nodes", C.c_uint32), ("bvh_depth", C.c_uint32),
                ("n_analytic_surfaces", C.c_uint32), ("n_textures", C.c_uint32), ("reserved", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("flatten_ms", C.c_double), ("upload_ms", C.c_double)]


_P = C.c_void_p
_FP = C.POINTER(C.c_float)
_UP = C.POINTER(C.c_uint32)
_U32, _I32, _U64, _F = C.c_uint32, C.c_int32, C.c_uint64, C.c_float

# name -> argtypes; every function returns int32 status except the two noted below
SIGNATURES = {
    "vr_context_create": [_I32, _P, C.POINTER(_P)],
    "vr_context_destroy": [_P],
    "vr_context_create_multi": [C.POINTER(C.c_int32), _U32, C.POINTER(_P)],
    "vr_context_device_count": [_P, _UP],
    "vr_scene_create": [_P, C.POINTER(_P)],
    "vr_scene_destroy": [_P],
    "vr_scene_add_texture_rgb32f": [_P, _FP, _U32, _U32, _I32, _UP],
    "vr_scene_add_mesh": [_P, _FP, _FP, _FP, _U32, _UP, _U32, _UP],
    "vr_scene_add_mesh_from_obj_file": [_P, C.c_char_p, _UP, _UP, _UP],
    "vr_scene_add_texture_rgb8": [_P, C.POINTER(C.c_uint8), _U32, _U32, _U32, _I32, _UP],
    "vr_scene_add_texture_rgb16": [_P, C.POINTER(C.c_uint16), _U32, _U32, _U32, _I32, _UP],
    "vr_scene_add_image_texture_file": [_P, C.c_char_p, _I32, _UP],
    "vr_image_load_rgb32f": [C.c_char_p, _UP, _UP, C.POINTER(_FP)],
    "vr_image_free": [_FP],
    "vr_scene_set_environment_hdri_file": [_P, C.c_char_p],
    "vr_debug_reference_leaf_order": [_FP, C.c_uint64, _UP],
    "vr_debug_flatten_mesh_digest": [_FP, _FP, _FP, C.c_uint32, _UP, C.c_uint32, C.POINTER(C.c_uint64), _UP, _UP,
                                     C.POINTER(C.c_double)],
    "vr_obj_load": [C.c_char_p, C.POINTER(ObjMeshC)],
    "vr_obj_free": [C.POINTER(ObjMeshC)],
    "vr_scene_add_sphere": [_P, _FP, _F, _UP],
    "vr_scene_add_ground_plane": [_P, _F, _UP],
    "vr_scene_add_material": [_P, C.POINTER(MaterialDescC), _UP],
    "vr_scene_add_object": [_P, _U32, _U32, _UP],
    "vr_scene_set_camera": [_P, _FP, _FP, _FP, _F, _I32, _F, _FP],
    "vr_scene_set_camera_look_at": [_P, _FP, _FP, _FP, _F],
    "vr_camera_look_at": [_FP, _FP, _FP, _FP, _FP],
    "vr_scene_set_environment_uniform": [_P, _FP],
    "vr_scene_set_environment_hdri_rgb32f": [_P, _FP, _U32, _U32],
    "vr_scene_clear_environment": [_P],
    "vr_scene_commit": [_P],
    "vr_scene_get_info": [_P, C.POINTER(SceneInfoC)],
    "vr_render_begin": [_P, _U32, _U32, C.POINTER(RenderSettingsC), C.POINTER(_P)],
    "vr_render_end": [_P],
    "vr_render_clear": [_P],
    "vr_render_accumulate": [_P, _U32],
    "vr_render_cancel": [_P],
    "vr_render_stats": [_P, C.POINTER(StatsC)],
    "vr_render_read_accum": [_P, _FP],
    "vr_render_accum_device_ptr": [_P, C.POINTER(_P)],
    "vr_render_resolve": [_P, _F, _F, _F, _I32, _FP],
    "vr_render_export_accum": [_P, C.POINTER(C.c_uint8)],
    "vr_render_reduce_peers": [_P, C.POINTER(C.c_uint8), _U32],
    "vr_render_resolve_peers": [_P, C.POINTER(C.c_uint8), _U32, _F, _F, _F, _I32, _FP],
    "vr_render_reduce_peer_ptrs": [_P, C.POINTER(C.c_void_p), _U32],
    "vr_render_resolve_peer_ptrs": [_P, C.POINTER(C.c_void_p), _U32, _F, _F, _F, _I32, _FP],
    "vr_debug_trace_primary": [_P, _U32, _UP, _UP, _FP],
    "vr_debug_trace_rays": [_P, _U64, _FP, _FP, _UP, _UP, _FP],
    "vr_debug_sample_radiance": [_P, _U64, _UP, _UP, _FP],
    "vr_debug_tie_ranks": [_P, _U32, _UP, _U32],
    "vr_debug_texture_sample": [_P, _U32, _U64, _FP, _FP],
    "vr_debug_environment_sample": [_P, _U64, _FP, _FP],
    "vr_debug_rng_draws": [_P, _U64, _U32, _U32, _U32, _UP],
    "vr_debug_unit_sphere": [_P, _U64, _U32, _U32, _U32, _FP],
}
NON_STATUS = {"vr_last_error": (C.c_char_p, []), "vr_abi_version": (C.c_uint32, [])}

_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built: build it with
    `python -c "import __graft_entry__ as g; g.build()"` or `make -C voidray_b200/csrc`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                          "(make -C voidray_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int32
    for name, (restype, argtypes) in NON_STATUS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == VR_OK:
        return
    msg = load().vr_last_error()
    msg = msg.decode("utf-8", "replace") if msg else ""
    if status == VR_ERR_CANCELLED:
        raise RenderCancelled(status, msg)
    raise VoidrayError(status, msg)


def fptr(a):
    return a.ctypes.data_as(_FP)


def uptr(a):
    return a.ctypes.data_as(_UP)
