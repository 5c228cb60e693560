"""Host-side mirror of the reference's scene-construction API.

Same names, argument meaning and handle semantics as the Rust builder, so scene recipes read like
the reference's `voidray_app/src/examples/*.rs`:

  Scene                      voidray_renderer/src/core/scene.rs:36-161
  Camera / Camera.look_at    voidray_renderer/src/core/camera.rs:7-36
  Materials.*                voidray_common/src/simple.rs:17-58
  Surfaces.*                 voidray_common/src/surfaces.rs:9-29
  Environments.*             voidray_common/src/environments.rs:9-17
  SampleType                 voidray_renderer/src/core/texture.rs:23-26
  Settings / RenderSettings / ColorManagementSettings / Tonemap / RenderMode
                             voidray_renderer/src/core/settings.rs

A `Scene` is pure host data (numpy arrays). `Scene.build_acceleration(ctx)` (voidray_b200.render)
flattens and uploads it through the C ABI, like `Accelerable::build_acceleration` (scene.rs:163-179).
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple, Union

import numpy as np

F32 = np.float32


def hex_color(x: int) -> Tuple[float, float, float]:
    """util/color.rs:18-23 — each channel is `(byte as f32) / 255.0`."""
    r = F32((x >> 16) & 0xFF) / F32(255.0)
    g = F32((x >> 8) & 0xFF) / F32(255.0)
    b = F32(x & 0xFF) / F32(255.0)
    return (float(r), float(g), float(b))


class SampleType(enum.IntEnum):
    Nearest = 0
    Bilinear = 1


class RenderMode(enum.IntEnum):
    Full = 0
    Normal = 1


class Tonemap(enum.IntEnum):  # settings.rs:36-55 (`as_i32`)
    NONE = 0
    ACES = 1
    Reinhard = 2
    Filmic = 3
    Uncharted2 = 4


class PixelMapping(enum.IntEnum):
    Fixed = 0      # y = index / width
    Reference = 1  # y = index / height with u32 wrapping, render/iterative.rs:26,33


@dataclass
class RenderSettings:  # settings.rs:15-33
    total_samples: int = 100
    update_frequency: float = 0.1
    render_mode: RenderMode = RenderMode.Full
    firefly_clamp: float = 3.0
    max_bounces: int = 10
    # not in the reference: the generator is seedable, the pixel mapping bug is switchable, and a
    # render may own a sub-range of the global sample indices (multi-GPU sample-range sharding)
    seed: int = 0x5EED0001
    pixel_mapping: PixelMapping = PixelMapping.Fixed
    sample_offset: int = 0
    max_paths_in_flight: int = 0
    # 0 = the reference estimator (parity); 1 = "fast": the same integrand sampled with one-sample MIS between the
    # reference's Lambertian density and an HDRI luminance table, plus Russian roulette from depth 3
    integrator: int = 0


@dataclass
class ColorManagementSettings:  # settings.rs:57-74
    tonemap: Tonemap = Tonemap.NONE
    gamma: float = 2.2
    exposure: float = 0.0
    transparent: bool = True


@dataclass
class Settings:  # settings.rs:3-7
    render: RenderSettings = field(default_factory=RenderSettings)
    color_management: ColorManagementSettings = field(default_factory=ColorManagementSettings)


# ---- materials ------------------------------------------------------------------------------------
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_EMISSION, MAT_LAMBERTIAN_BSDF, MAT_MICROFACET = range(6)


@dataclass
class MaterialDesc:
    kind: int
    color: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    param: float = 0.0
    albedo_tex: int = -1
    normal_tex: int = -1
    # MicrofacetBSDF fields (voidray_common/src/microfacet.rs:9-27)
    index: float = 1.5
    roughness: float = 1.0
    metallic: float = 0.0
    emittance: float = 0.0
    transparent: bool = False


class Materials:
    """voidray_common/src/simple.rs:17-58"""

    @staticmethod
    def lambertian(albedo) -> MaterialDesc:
        return MaterialDesc(MAT_LAMBERTIAN, tuple(albedo))

    @staticmethod
    def lambertian_bsdf(albedo) -> MaterialDesc:
        return MaterialDesc(MAT_LAMBERTIAN_BSDF, tuple(albedo))

    @staticmethod
    def lambertian_texture_no_normal(albedo: int) -> MaterialDesc:
        return MaterialDesc(MAT_LAMBERTIAN, albedo_tex=int(albedo))

    @staticmethod
    def lambertian_texture(albedo: int, normal: int) -> MaterialDesc:
        return MaterialDesc(MAT_LAMBERTIAN, albedo_tex=int(albedo), normal_tex=int(normal))

    @staticmethod
    def metal(albedo, fuzz: float) -> MaterialDesc:
        return MaterialDesc(MAT_METAL, tuple(albedo), float(fuzz))

    @staticmethod
    def dielectric(ir: float) -> MaterialDesc:
        return MaterialDesc(MAT_DIELECTRIC, (0.0, 0.0, 0.0), float(ir))

    @staticmethod
    def emissive(strength: float) -> MaterialDesc:
        return MaterialDesc(MAT_EMISSION, (1.0, 1.0, 1.0), float(strength))

    @staticmethod
    def colored_emissive(color, strength: float) -> MaterialDesc:
        return MaterialDesc(MAT_EMISSION, tuple(color), float(strength))


class MicrofacetBSDF:
    """voidray_common/src/microfacet.rs:29-101 (constructors of the Beckmann / Cook-Torrance BSDF)."""

    @staticmethod
    def _mk(color, index, roughness, metallic=0.0, emittance=0.0, transparent=False) -> MaterialDesc:
        return MaterialDesc(MAT_MICROFACET, tuple(color), index=float(index), roughness=float(roughness),
                            metallic=float(metallic), emittance=float(emittance), transparent=bool(transparent))

    @staticmethod
    def diffuse(color) -> MaterialDesc:
        return MicrofacetBSDF._mk(color, 1.5, 1.0)

    @staticmethod
    def specular(color, roughness: float) -> MaterialDesc:
        return MicrofacetBSDF._mk(color, 1.5, roughness)

    @staticmethod
    def clear(index: float, roughness: float) -> MaterialDesc:
        return MicrofacetBSDF._mk(hex_color(0xFFFFFF), index, roughness, transparent=True)

    @staticmethod
    def transparent(color, index: float, roughness: float) -> MaterialDesc:
        return MicrofacetBSDF._mk(color, index, roughness, transparent=True)

    @staticmethod
    def metallic(color, roughness: float) -> MaterialDesc:
        return MicrofacetBSDF._mk(color, 1.5, roughness, metallic=1.0)

    @staticmethod
    def light(color, emittance: float) -> MaterialDesc:
        return MicrofacetBSDF._mk(color, 1.0, 1.0, emittance=emittance)


# ---- surfaces -------------------------------------------------------------------------------------
@dataclass
class SphereDesc:
    center: Tuple[float, float, float]
    radius: float


@dataclass
class GroundPlaneDesc:
    height: float


@dataclass
class MeshData:
    """core/mesh.rs:35-41 before acceleration: the vertex buffer (position, uv, normal) and indices."""
    positions: np.ndarray  # (n, 3) f32
    uvs: np.ndarray        # (n, 2) f32
    normals: np.ndarray    # (n, 3) f32
    indices: np.ndarray    # (3 * n_triangles,) u32

    @staticmethod
    def from_buffers(positions, indices, uvs=None, normals=None) -> "MeshData":
        positions = np.ascontiguousarray(positions, dtype=F32).reshape(-1, 3)
        n = positions.shape[0]
        uvs = np.zeros((n, 2), F32) if uvs is None else np.ascontiguousarray(uvs, dtype=F32).reshape(-1, 2)
        normals = np.zeros((n, 3), F32) if normals is None else np.ascontiguousarray(normals, dtype=F32).reshape(-1, 3)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        return MeshData(positions, uvs, normals, indices)

    @staticmethod
    def from_file(path: str) -> "MeshData":
        """Mesh::from_file, core/mesh.rs:46-74 (obj-rs 0.7.0 `load_obj::<TexturedVertex, u32>`)."""
        from .assets import load_obj
        return load_obj(path)

    @property
    def n_triangles(self) -> int:
        return self.indices.size // 3


@dataclass
class ObjFile:
    """A mesh surface that stays a path until the scene is committed (Mesh::from_file, core/mesh.rs:46-74)."""
    path: str
    _mesh: Optional["MeshData"] = None

    def mesh(self) -> "MeshData":
        if self._mesh is None:
            self._mesh = MeshData.from_file(self.path)
        return self._mesh

    @property
    def n_triangles(self) -> int:
        return self.mesh().n_triangles


class Surfaces:
    """voidray_common/src/surfaces.rs:9-29"""

    @staticmethod
    def sphere(center, radius: float) -> SphereDesc:
        return SphereDesc(tuple(float(c) for c in center), float(radius))

    @staticmethod
    def ground_plane(height: float) -> GroundPlaneDesc:
        return GroundPlaneDesc(float(height))

    @staticmethod
    def quad(q1, q2, q3, q4) -> MeshData:
        # surfaces.rs:18-29: four position-only vertices, indices [0,1,2, 2,0,3]
        return MeshData.from_buffers(np.array([q1, q2, q3, q4], F32), np.array([0, 1, 2, 2, 0, 3], np.uint32))


# ---- environments ---------------------------------------------------------------------------------
@dataclass
class UniformEnvironment:
    color: Tuple[float, float, float]


@dataclass
class HDRIEnvironment:
    image: np.ndarray  # (h, w, 3) f32, row 0 first


class Environments:
    """voidray_common/src/environments.rs:9-17 — returns Option<Arc<dyn Environment>>."""

    @staticmethod
    def uniform(background) -> UniformEnvironment:
        return UniformEnvironment(tuple(float(c) for c in background))

    @staticmethod
    def hdri(path_or_image) -> HDRIEnvironment:
        if isinstance(path_or_image, str):  # environments.rs:42-55, decoded by the library (image_io.cpp)
            from .assets import load_image_native
            return HDRIEnvironment(load_image_native(path_or_image))
        return HDRIEnvironment(np.ascontiguousarray(path_or_image, dtype=F32))


# ---- camera ---------------------------------------------------------------------------------------
def _normalize(v: np.ndarray) -> np.ndarray:
    # cgmath: v * (1 / sqrt((x*x + y*y) + z*z)), all in f32
    m2 = F32(F32(v[0] * v[0]) + F32(v[1] * v[1])) + F32(v[2] * v[2])
    return (v * (F32(1.0) / np.sqrt(F32(m2), dtype=F32))).astype(F32)


@dataclass
class Camera:  # core/camera.rs:7-22
    eye: Tuple[float, float, float]
    direction: Tuple[float, float, float]
    up: Tuple[float, float, float]
    fov: float
    dof: Optional[Tuple[float, Tuple[float, float, float]]] = None

    @staticmethod
    def look_at(eye, center, up, fov: float) -> "Camera":
        """camera.rs:26-36 evaluated in f32 with cgmath's operation order."""
        e = np.array(eye, F32)
        c = np.array(center, F32)
        u0 = np.array(up, F32)
        direction = _normalize((c - e).astype(F32))
        d = F32(F32(u0[0] * direction[0]) + F32(u0[1] * direction[1])) + F32(u0[2] * direction[2])
        u = _normalize((u0 - (F32(d) * direction).astype(F32)).astype(F32))
        return Camera(tuple(float(x) for x in e), tuple(float(x) for x in direction), tuple(float(x) for x in u),
                      float(F32(fov)))


# ---- scene ----------------------------------------------------------------------------------------
@dataclass
class ImageTexture:
    image: np.ndarray  # (h, w, 3) f32
    sample_type: SampleType


@dataclass
class Object:  # scene.rs:31-34
    surface: int
    material: int


SurfaceDesc = Union[MeshData, "ObjFile", SphereDesc, GroundPlaneDesc]


class Scene:
    """core/scene.rs:36-161. Handles are plain ints (the reference's usize newtypes)."""

    def __init__(self) -> None:
        # Scene::empty(), scene.rs:95-111
        self.camera: Camera = Camera.look_at((1.0, 0.0, 10.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), float(F32(math.pi) / F32(6.0)))
        self.textures: List[ImageTexture] = []
        self.objects: List[Object] = []
        self.surfaces: List[SurfaceDesc] = []
        self.materials: List[MaterialDesc] = []
        self.environment: Optional[Union[UniformEnvironment, HDRIEnvironment]] = None

    @staticmethod
    def empty() -> "Scene":
        return Scene()

    def add_material(self, material: MaterialDesc) -> int:
        self.materials.append(material)
        return len(self.materials) - 1

    def add_analytic_surface(self, analytic: Union[SphereDesc, GroundPlaneDesc]) -> int:
        self.surfaces.append(analytic)
        return len(self.surfaces) - 1

    def add_mesh(self, mesh: MeshData) -> int:
        self.surfaces.append(mesh)
        return len(self.surfaces) - 1

    def add_mesh_from_file(self, path: str) -> int:
        """scene.rs:137-144. The OBJ file is parsed by the library's native loader at build_acceleration
        (vr_scene_add_mesh_from_obj_file); `assets.load_obj` is the same loader in Python for host-side use."""
        self.surfaces.append(ObjFile(path))
        return len(self.surfaces) - 1

    def add_object(self, material: int, surface: int) -> int:
        self.objects.append(Object(surface=int(surface), material=int(material)))
        return len(self.objects) - 1

    def add_image_texture(self, path_or_image, sample_type: SampleType) -> int:
        if isinstance(path_or_image, str):  # texture.rs:36-49, decoded by the library (image_io.cpp)
            from .assets import load_image_native
            image = load_image_native(path_or_image)
        else:
            image = np.ascontiguousarray(path_or_image, dtype=F32)
        assert image.ndim == 3 and image.shape[2] == 3
        self.textures.append(ImageTexture(image, SampleType(sample_type)))
        return len(self.textures) - 1

    def build_acceleration(self, ctx=None):
        """Accelerable::build_acceleration (scene.rs:163-179) -> SceneAcceleration on the device."""
        from .render import SceneAcceleration
        return SceneAcceleration(self, ctx)

    # convenience for statistics
    def n_triangles(self) -> int:
        return sum(s.n_triangles for s in self.surfaces if isinstance(s, (MeshData, ObjFile)))
