"""voidray_b200 — B200-native (sm_100a CUDA) implementation of voidray's progressive path-tracing
hot path behind the reference renderer's API. See DESIGN.md.

`scene`/`assets`/`scenes` are pure host code. `render` binds libvoidray_cuda.so through ctypes and
raises if the library is missing — there is no CPU fallback."""
from .scene import (Camera, ColorManagementSettings, Environments, Materials, MeshData, MicrofacetBSDF, PixelMapping,
                    RenderMode,
                    RenderSettings, SampleType, Scene, Settings, Surfaces, Tonemap, hex_color)

__all__ = ["Camera", "ColorManagementSettings", "Environments", "Materials", "MeshData", "MicrofacetBSDF", "PixelMapping",
           "RenderMode", "RenderSettings", "SampleType", "Scene", "Settings", "Surfaces", "Tonemap", "hex_color"]
