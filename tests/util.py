"""Shared helpers for the parity tests."""
import numpy as np

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Camera, Environments, Materials, Scene

F32 = np.float32
MISS = 0xFFFFFFFF


def single_mesh_scene(mesh, material=None, env=(0.5, 0.5, 0.5)):
    s = Scene.empty()
    sf = s.add_mesh(mesh)
    s.add_object(s.add_material(material or Materials.lambertian((0.5, 0.5, 0.5))), sf)
    if env is not None:
        s.environment = Environments.uniform(env)
    return s


def obj_scene(name, **kw):
    return single_mesh_scene(load_obj(asset_path(name)), **kw)


def scene_bounds(scene):
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for s in scene.surfaces:
        if hasattr(s, "positions"):
            lo = np.minimum(lo, s.positions.min(0))
            hi = np.maximum(hi, s.positions.max(0))
        elif hasattr(s, "center"):
            lo = np.minimum(lo, np.array(s.center) - s.radius)
            hi = np.maximum(hi, np.array(s.center) + s.radius)
    return lo, hi


def random_rays(n, lo, hi, seed):
    """Incoherent rays: origins on a shell around the box, aimed at random points inside it."""
    rng = np.random.default_rng(seed)
    c = (lo + hi) / 2
    r = np.linalg.norm(hi - lo) / 2
    o = rng.normal(size=(n, 3))
    o /= np.linalg.norm(o, axis=1, keepdims=True)
    o = c + o * r * rng.uniform(0.2, 1.5, (n, 1))
    tgt = c + rng.uniform(-1, 1, (n, 3)) * r * 0.6
    return o.astype(F32), (tgt - o).astype(F32)


def rel_mse(a, b):
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def rmse(a, b):
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))
