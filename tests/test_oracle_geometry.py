"""Oracle self-checks for the closest-hit path (core/scene.rs:182-185, bvh.rs, mesh.rs, surfaces.rs):
closed-form distances, agreement of the three traversal modes, and the tie rule."""
import numpy as np

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Materials, MeshData, Scene, Surfaces

from util import MISS, F32, obj_scene, random_rays, scene_bounds, single_mesh_scene


def test_single_triangle_distance(oracle):
    tri = MeshData.from_buffers(np.array([[-1, -1, 2], [1, -1, 2], [0, 1, 2]], F32), [0, 1, 2])
    osc = oracle.OracleScene(single_mesh_scene(tri))
    o = np.array([[0, 0, 0], [0, 0, 0], [0, 0, 5], [0, 0, 2.000001]], F32)
    d = np.array([[0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, 1]], F32)
    s, p, t, _ = osc.trace_rays(o, d)
    assert s[0] == 0 and p[0] == 0 and t[0] == F32(2.0)
    assert s[1] == MISS                      # behind the origin
    assert p[2] == 0 and t[2] == F32(3.0)    # no back-face culling (mesh.rs:156)
    assert s[3] == MISS                      # nothing in front


def test_cube_axis_distances(oracle):
    osc = oracle.OracleScene(obj_scene("cube.obj"))  # [-1, 1]^3
    o = np.array([[5, 0.3, 0.2], [0.1, -7, 0.4], [0.2, 0.1, 3], [0.3, 0.2, 0.1]], F32)
    d = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], F32)
    _, _, t, _ = osc.trace_rays(o, d)
    assert np.allclose(t, [4.0, 6.0, 2.0, 0.9], rtol=1e-6)


def test_small_mesh_uses_linear_loop_and_first_wins(oracle):
    # two coincident triangles in a <= 4-triangle mesh: the first index wins the tie (mesh.rs:131)
    v = np.array([[-1, -1, 2], [1, -1, 2], [0, 1, 2]], F32)
    m = MeshData.from_buffers(np.concatenate([v, v]), [0, 1, 2, 3, 4, 5])
    osc = oracle.OracleScene(single_mesh_scene(m))
    for mode in (oracle.MODE_FAITHFUL, oracle.MODE_BRUTE):
        _, p, t, _ = osc.trace_rays(np.zeros((1, 3), F32), np.array([[0, 0, 1]], F32), mode)
        assert p[0] == 0 and t[0] == F32(2.0)
    assert list(osc.mesh_tie_rank(0)) == [1, 0]


def test_bvh_mesh_right_wins_ties(oracle):
    # six triangles (> SMALL_MESH) with 0 and 5 coincident: the reference returns the one that is
    # later in its tree's in-order leaf sequence (bvh.rs:171)
    base = np.array([[-1, -1, 2], [1, -1, 2], [0, 1, 2]], F32)
    vs = [base] + [base + np.array([3.0 * k, 0, 0], F32) for k in range(1, 5)] + [base]
    m = MeshData.from_buffers(np.concatenate(vs), np.arange(18))
    osc = oracle.OracleScene(single_mesh_scene(m))
    rank = osc.mesh_tie_rank(0)
    want = 0 if rank[0] > rank[5] else 5
    for mode in (oracle.MODE_FAITHFUL, oracle.MODE_BRUTE):
        _, p, _, _ = osc.trace_rays(np.zeros((1, 3), F32), np.array([[0, 0, 1]], F32), mode)
        assert p[0] == want
    assert sorted(rank) == list(range(6))


def test_traversal_modes_agree(oracle):
    for name, n in (("cube.obj", 20000), ("mushroom.obj", 20000), ("mossy_ground.obj", 5000)):
        scene = obj_scene(name)
        osc = oracle.OracleScene(scene)
        o, d = random_rays(n, *scene_bounds(scene), seed=11)
        sF, pF, tF, cF = osc.trace_rays(o, d, oracle.MODE_FAITHFUL)
        sE, pE, tE, cE = osc.trace_rays(o, d, oracle.MODE_EARLY_OUT)
        assert np.array_equal(pF, pE) and np.array_equal(tF, tE)
        assert cE.box_tests <= cF.box_tests and cE.tri_tests <= cF.tri_tests
        nb = min(n, 3000)
        sB, pB, tB, _ = osc.trace_rays(o[:nb], d[:nb], oracle.MODE_BRUTE)
        assert np.array_equal(pF[:nb], pB) and np.array_equal(tF[:nb], tB) and np.array_equal(sF[:nb], sB)
        assert 0.05 < (sF != MISS).mean() < 1.0


def test_sphere_and_plane(oracle):
    s = Scene.empty()
    sp = s.add_analytic_surface(Surfaces.sphere((0, 0, 5), 1.0))
    gp = s.add_analytic_surface(Surfaces.ground_plane(-2.0))
    mat = s.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    s.add_object(mat, sp)
    s.add_object(mat, gp)
    osc = oracle.OracleScene(s)
    o = np.array([[0, 0, 0], [0, 0, 5], [0, 0, 0], [0, 5, 0]], F32)
    d = np.array([[0, 0, 1], [0, 0, 1], [0, -1, 0], [0, 1, 0]], F32)
    sf, _, t, n, uv, front, _ = osc.trace_rays(o, d, details=True)
    assert sf[0] == 0 and t[0] == F32(4.0) and front[0] == 1 and np.allclose(n[0], [0, 0, -1])
    assert sf[1] == 0 and t[1] == F32(1.0) and front[1] == 0   # from inside: far root, flipped normal
    assert np.allclose(n[1], [0, 0, -1])
    assert sf[2] == 1 and t[2] == F32(2.0) and np.allclose(uv[2], [0, 0])
    assert sf[3] == MISS


def test_hit_record_details_on_cube(oracle):
    # cube.obj has smooth corner normals (+-0.5773 each). Near a corner the interpolated normal is more
    # than 30 degrees from the face normal, so the geometric normal replaces it (mesh.rs:179-181);
    # towards the face centre the interpolated normal is kept *un-normalised* (mesh.rs:176).
    osc = oracle.OracleScene(obj_scene("cube.obj"))
    o = np.array([[0.97, 0.97, 5], [0.25, 0.5, 5]], F32)
    d = np.array([[0, 0, -1], [0, 0, -1]], F32)
    sf, p, t, n, uv, front, _ = osc.trace_rays(o, d, details=True)
    assert np.all(t == F32(4.0)) and np.all(front == 1)
    assert np.allclose(n[0], [0, 0, 1], atol=1e-6)
    ln = np.linalg.norm(n[1])
    assert 0.3 < ln < 0.99 and n[1, 2] > 0            # interpolated, not unit length
    ang = np.degrees(np.arctan2(np.linalg.norm(np.cross(n[1], [0, 0, 1])), n[1, 2]))
    assert ang <= 30.0
    assert np.all((uv >= 0.0) & (uv <= 1.0))
