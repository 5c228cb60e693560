"""Test infrastructure: writes a voidray_b200 Scene in the flat binary form tests/c/scene_file.h reads, mirroring call
for call what SceneAcceleration.__init__ (voidray_b200/render.py) sends through the C ABI."""
import struct

import numpy as np

from voidray_b200.scene import (GroundPlaneDesc, HDRIEnvironment, MeshData, ObjFile, SphereDesc, UniformEnvironment)

F32 = np.float32


def write_scene(path, scene):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 0x56525343))
        f.write(struct.pack("<I", len(scene.textures)))
        for tex in scene.textures:
            img = np.ascontiguousarray(tex.image, dtype=F32)
            f.write(struct.pack("<IIi", img.shape[1], img.shape[0], int(tex.sample_type)))
            f.write(img.tobytes())
        f.write(struct.pack("<I", len(scene.surfaces)))
        for surf in scene.surfaces:
            if isinstance(surf, ObjFile):
                surf = surf.mesh()
            if isinstance(surf, MeshData):
                pos = np.ascontiguousarray(surf.positions, dtype=F32)
                uvs = np.ascontiguousarray(surf.uvs, dtype=F32)
                nrm = np.ascontiguousarray(surf.normals, dtype=F32)
                idx = np.ascontiguousarray(surf.indices, dtype=np.uint32).ravel()
                f.write(struct.pack("<III", 0, pos.shape[0], idx.size))
                f.write(pos.tobytes() + uvs.tobytes() + nrm.tobytes() + idx.tobytes())
            elif isinstance(surf, SphereDesc):
                f.write(struct.pack("<I4f", 1, *[float(F32(x)) for x in surf.center], float(F32(surf.radius))))
            elif isinstance(surf, GroundPlaneDesc):
                f.write(struct.pack("<If", 2, float(F32(surf.height))))
            else:
                raise TypeError(surf)
        f.write(struct.pack("<I", len(scene.materials)))
        for m in scene.materials:
            f.write(struct.pack("<i3ffii4fi", int(m.kind), *[float(F32(c)) for c in m.color], float(F32(m.param)),
                                int(m.albedo_tex), int(m.normal_tex), float(F32(m.index)), float(F32(m.roughness)),
                                float(F32(m.metallic)), float(F32(m.emittance)), 1 if m.transparent else 0))
        f.write(struct.pack("<I", len(scene.objects)))
        for o in scene.objects:
            f.write(struct.pack("<II", o.material, o.surface))
        cam = scene.camera
        dof = cam.dof
        f.write(struct.pack("<10fi4f", *[float(F32(x)) for x in cam.eye], *[float(F32(x)) for x in cam.direction],
                            *[float(F32(x)) for x in cam.up], float(F32(cam.fov)), 1 if dof else 0,
                            float(F32(dof[0])) if dof else 0.0, *([float(F32(x)) for x in dof[1]] if dof else [0.0, 0.0, 0.0])))
        env = scene.environment
        if isinstance(env, UniformEnvironment):
            f.write(struct.pack("<i3f", 1, *[float(F32(c)) for c in env.color]))
        elif isinstance(env, HDRIEnvironment):
            img = np.ascontiguousarray(env.image, dtype=F32)
            f.write(struct.pack("<iII", 2, img.shape[1], img.shape[0]))
            f.write(img.tobytes())
        else:
            f.write(struct.pack("<i", 0))
