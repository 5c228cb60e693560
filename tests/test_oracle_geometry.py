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


def test_thirty_degree_normal_threshold(oracle):
    # core/mesh.rs:176-181: the interpolated normal is replaced by the geometric one when
    # cgmath 0.18 Vector3::angle(n, n_geo) = atan2(|n x n_geo|, n . n_geo) exceeds degrees_to_radians(30) = 30 * PI / 180
    # (util/math.rs:32-34), strictly. Known answers around the threshold: vertex normals tilted from the geometric normal
    # (0, 0, -1 for this winding seen from -z ... the hit record flips it towards the ray) by a known angle in the xz plane.
    from voidray_b200.scene import MeshData
    base = np.array([[-1, -1, 2], [1, -1, 2], [0, 1, 2]], F32)
    thr = F32(30.0) * F32(np.pi) / F32(180.0)
    flipped = []
    for deg in (0.0, 10.0, 29.0, 29.99, 29.999, 30.001, 30.01, 31.0, 60.0, 89.0):
        a = np.deg2rad(deg)
        # geometric normal of this triangle: ((v2 - v1) x (v0 - v1)) normalised (mesh.rs:80-84)
        ng = np.cross(base[2] - base[1], base[0] - base[1]).astype(F32)
        ng = ng / np.linalg.norm(ng)
        tilt = np.array([np.sin(a), 0.0, np.cos(a) * ng[2]], np.float64).astype(F32)  # same hemisphere as ng, tilted in x
        mesh = MeshData.from_buffers(base, [0, 1, 2], normals=np.tile(tilt, (3, 1)))
        osc = oracle.OracleScene(single_mesh_scene(mesh))
        s, p, t, normal, uv, front, _ = osc.trace_rays(np.array([[0.1, -0.2, 0]], F32), np.array([[0, 0, 1]], F32), details=True)
        assert s[0] == 0 and t[0] == F32(2.0)
        # the oracle's decision against an independent f32 evaluation of the same expression
        cr = np.cross(tilt.astype(F32), ng.astype(F32)).astype(F32)
        mag = np.sqrt(((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]).astype(F32)).astype(F32)
        dt = ((tilt[0] * ng[0] + tilt[1] * ng[1]) + tilt[2] * ng[2]).astype(F32)
        ang = np.arctan2(mag, dt).astype(F32)
        is_geo = bool(np.allclose(np.abs(normal[0]), np.abs(ng), atol=1e-6))
        is_tilt = bool(np.allclose(np.abs(normal[0]), np.abs(tilt), atol=1e-6))  # (u + v + w) * tilt, a rounding off
        if abs(float(ang) - float(thr)) > 1e-5:  # away from the last ulps of atan2 the answer is known
            assert (is_geo if ang > thr else is_tilt), (deg, ang, thr, normal[0])
        flipped.append(is_geo and not is_tilt)
    assert flipped[1:3] == [False, False] and flipped[-3:] == [True, True, True]
