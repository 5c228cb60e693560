"""Shared helpers for the parity tests."""
import numpy as np

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Camera, Environments, Materials, Scene, Surfaces

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


def write_obj(path, positions, faces):
    with open(path, "w") as f:
        for p in positions:
            f.write(f"v {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
        f.write("vt 0 0\nvn 0 0 1\n")
        for a, b, c in faces:
            f.write(f"f {a + 1}/1/1 {b + 1}/1/1 {c + 1}/1/1\n")


# ---- scenes whose reference scene-level tree has a flat Split (core/scene.rs:182-185, util/aabb.rs:86-148) ----
def quad_obj(path, corners):
    """A quad as the two triangles of Surfaces::quad (voidray_common/src/surfaces.rs:18-28): (0, 1, 2), (0, 2, 3)."""
    write_obj(path, np.array(corners, F32), [(0, 1, 2), (0, 2, 3)])


def flat_split_cases(tmp_path):
    """name -> (surfaces as harness arguments, the same scene for the oracle). Every case has a scene-level Split whose
    box is flat on an axis, which the reference rejects for every ray with a component along it."""
    def quad_y0(x0, x1, tag):
        p = str(tmp_path / f"quad_{tag}.obj")
        quad_obj(p, [(x0, 0, 0), (x0, 0, 1), (x1, 0, 1), (x1, 0, 0)])
        return p
    cases = {}
    tiles = [quad_y0(2.0 * k, 2.0 * k + 1.0, f"t{k}") for k in range(5)]
    cases["two_coplanar_quads"] = [("obj", tiles[0]), ("obj", tiles[1])]
    cases["three_coplanar_quads"] = [("obj", tiles[0]), ("obj", tiles[1]), ("obj", tiles[2])]
    cases["five_coplanar_quads"] = [("obj", t) for t in tiles]
    # a flat mesh and a sphere: the root Split is not flat, nothing may be culled that the reference keeps
    cases["flat_mesh_and_sphere"] = [("obj", tiles[0]), ("sphere", 0.5, 0.75, 0.5, 0.5)]
    # a flat two-surface sub-list inside a larger scene: mushroom | (quad, quad) — only the pair is culled
    far = [str(tmp_path / "far_a.obj"), str(tmp_path / "far_b.obj")]
    quad_obj(far[0], [(20, 0.5, 0), (20, 0.5, 1), (21, 0.5, 1), (21, 0.5, 0)])
    quad_obj(far[1], [(22, 0.5, 0), (22, 0.5, 1), (23, 0.5, 1), (23, 0.5, 0)])
    cases["flat_pair_inside_larger_scene"] = [("obj", asset_path("mushroom.obj")), ("obj", far[0]), ("obj", far[1]),
                                              ("sphere", 10.0, 1.0, 0.0, 1.0)]
    # walls in the x = const and z = const planes as well
    wx = [str(tmp_path / "wx_a.obj"), str(tmp_path / "wx_b.obj")]
    quad_obj(wx[0], [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)])
    quad_obj(wx[1], [(1, 2, 0), (1, 3, 0), (1, 3, 1), (1, 2, 1)])
    cases["two_coplanar_walls_x"] = [("obj", wx[0]), ("obj", wx[1])]
    return cases


def build_case_scene(surfaces):
    scene = Scene.empty()
    mat = scene.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    args = []
    for s in surfaces:
        if s[0] == "obj":
            scene.add_object(mat, scene.add_mesh(load_obj(s[1])))
            args += ["obj", s[1]]
        else:
            scene.add_object(mat, scene.add_analytic_surface(Surfaces.sphere(tuple(s[1:4]), s[4])))
            args += ["sphere", *[repr(float(v)) for v in s[1:]]]
    scene.environment = Environments.uniform((0.5, 0.5, 0.5))
    return scene, args


def flat_case_rays(scene, seed):
    """Rays aimed at points on the surfaces (so that the un-culled answer is a hit): oblique, axis-parallel along each
    axis (zero components: the reference's 1 / 0 slabs), and starting on a surface."""
    rng = np.random.default_rng(seed)
    targets = []
    for s in scene.surfaces:
        if hasattr(s, "positions"):
            tri = s.positions[s.indices.reshape(-1, 3)[rng.integers(0, len(s.indices) // 3, 400)]]
            w = rng.dirichlet((1, 1, 1), 400)
            targets.append((tri * w[:, :, None]).sum(1))
        else:
            v = rng.normal(size=(400, 3))
            targets.append(np.array(s.center) + s.radius * v / np.linalg.norm(v, axis=1, keepdims=True))
    tgt = np.concatenate(targets)
    o = tgt + rng.normal(size=tgt.shape) * 2.0
    d = tgt - o
    k = len(tgt) // 4
    for axis in range(3):  # axis-parallel rays through the targets
        sel = slice(axis * (k // 3), (axis + 1) * (k // 3))
        o[sel] = tgt[sel]
        o[sel, axis] += 3.0
        d[sel] = 0.0
        d[sel, axis] = -1.0
    sel = slice(k, k + k // 2)  # one zero component
    d[sel, rng.integers(0, 3)] = 0.0
    return o.astype(F32), d.astype(F32)
