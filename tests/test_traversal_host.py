"""The closest-hit traversal source of the CUDA kernels (voidray_b200/csrc/traversal.cuh) compiled for the CPU
(tests/c/trav_host.cpp + tests/c/host_shim.h) and compared with the oracle bit for bit: the k_trace_rays gate of
tests/test_gpu_closest_hit.py without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from voidray_b200.assets import asset_path, load_obj
from voidray_b200.scene import Environments, Materials, Scene, Surfaces

from util import (F32, MISS, build_case_scene, flat_case_rays, flat_split_cases, quad_obj, random_rays, scene_bounds,
                  write_obj)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {"default": []}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    built = {}

    def get(variant):
        if variant not in built:
            exe = str(tmp_path_factory.mktemp("trav") / f"trav_host_{variant}")
            csrc = os.path.join(ROOT, "voidray_b200", "csrc")
            r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-ffp-contract=off", "-DVR_HOST_SHIM",
                                *VARIANTS[variant], "-I", os.path.join(ROOT, "tests", "c"), "-I", csrc, "-x", "c++",
                                os.path.join(csrc, "scene_build.cpp"), os.path.join(ROOT, "tests", "c", "trav_host.cpp"),
                                "-o", exe], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            built[variant] = exe
        return built[variant]
    return get


def run_harness(exe, tmp_path, origins, dirs, surfaces):
    rays = np.concatenate([origins, dirs], axis=1).astype(F32)
    rp, op = str(tmp_path / "rays.bin"), str(tmp_path / "out.bin")
    rays.tofile(rp)
    r = subprocess.run([exe, rp, op, *surfaces], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    out = np.fromfile(op, dtype=np.dtype([("surface", "<u4"), ("prim", "<u4"), ("t", "<f4")]))
    assert len(out) == len(rays)
    return out, r.stdout


def mixed_rays(scene, n, seed):
    lo, hi = scene_bounds(scene)
    o, d = random_rays(n, lo, hi, seed)
    # a share of axis-parallel and exactly diagonal directions (zero components, equal slab distances)
    rng = np.random.default_rng(seed + 1)
    k = n // 8
    d[:k, rng.integers(0, 3, k)] = 0.0
    d[k:2 * k] = np.sign(d[k:2 * k]) + (d[k:2 * k] == 0)
    return o, d


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name,n", [("cube.obj", 20000), ("mushroom.obj", 40000), ("mossy_ground.obj", 30000)])
def test_kernel_traversal_source_matches_oracle(oracle, harness, tmp_path, variant, name, n):
    scene = Scene.empty()
    scene.add_object(scene.add_material(Materials.lambertian((0.5, 0.5, 0.5))), scene.add_mesh(load_obj(asset_path(name))))
    scene.environment = Environments.uniform((0.5, 0.5, 0.5))
    o, d = mixed_rays(scene, n, 11)
    s_ref, p_ref, t_ref, _ = oracle.OracleScene(scene).trace_rays(o, d)
    out, log = run_harness(harness(variant), tmp_path, o, d, ["obj", asset_path(name)])
    assert (s_ref != MISS).mean() > 0.1
    assert np.array_equal(out["surface"], s_ref) and np.array_equal(out["prim"], p_ref)
    assert np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))  # bit-equal distances, inf on a miss


@pytest.mark.parametrize("variant", ["default"])
def test_kernel_traversal_source_two_meshes_and_analytic_surfaces(oracle, harness, tmp_path, variant):
    # two meshes, a sphere inside the scene and a ground plane: surface handles, the analytic pass after the BVH and
    # the tie ranks across surfaces
    scene = Scene.empty()
    mat = scene.add_material(Materials.lambertian((0.5, 0.5, 0.5)))
    scene.add_object(mat, scene.add_mesh(load_obj(asset_path("mushroom.obj"))))
    scene.add_object(mat, scene.add_analytic_surface(Surfaces.sphere((0.5, 1.0, 0.25), 0.75)))
    scene.add_object(mat, scene.add_mesh(load_obj(asset_path("fancy_monkey.obj"))))
    scene.add_object(mat, scene.add_analytic_surface(Surfaces.ground_plane(-0.125)))
    scene.environment = Environments.uniform((0.5, 0.5, 0.5))
    o, d = mixed_rays(scene, 30000, 5)
    s_ref, p_ref, t_ref, _ = oracle.OracleScene(scene).trace_rays(o, d)
    out, _ = run_harness(harness(variant), tmp_path, o, d,
                         ["obj", asset_path("mushroom.obj"), "sphere", "0.5", "1.0", "0.25", "0.75",
                          "obj", asset_path("fancy_monkey.obj"), "plane", "-0.125"])
    assert len(set(s_ref.tolist())) >= 4
    assert np.array_equal(out["surface"], s_ref)
    tri = p_ref != MISS
    assert np.array_equal(out["prim"][tri], p_ref[tri])
    assert np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))


@pytest.mark.parametrize("variant", ["default"])
def test_kernel_traversal_source_tie_rule_and_surface_starts(oracle, harness, tmp_path, variant):
    # (1) three coincident layers of an 8 x 8 quad grid (shared edges and vertices, every hit distance tied three to
    # eighteen ways): the winner is the reference's right-most leaf (bvh.rs:171), whatever order the traversal takes;
    # rays through vertices, edge mid-points and interiors, axis-parallel
    n = 8
    xs = np.arange(n + 1, dtype=F32) * F32(0.25) - F32(1.0)
    grid = np.array([[x, y, 2.0] for y in xs for x in xs], F32)
    quads = []
    for j in range(n):
        for i in range(n):
            a, b, c, d = j * (n + 1) + i, j * (n + 1) + i + 1, (j + 1) * (n + 1) + i + 1, (j + 1) * (n + 1) + i
            quads += [(a, b, c), (a, c, d)]
    pos = np.concatenate([grid, grid, grid])
    faces = [(a + k * len(grid), b + k * len(grid), c + k * len(grid)) for k in range(3) for a, b, c in quads]
    obj = str(tmp_path / "layers.obj")
    write_obj(obj, pos, faces)
    scene = Scene.empty()
    scene.add_object(scene.add_material(Materials.lambertian((0.5, 0.5, 0.5))), scene.add_mesh(load_obj(obj)))
    scene.environment = Environments.uniform((0.5, 0.5, 0.5))
    pts = np.arange(4 * n + 1, dtype=F32) * F32(0.0625) - F32(1.0)  # vertices, edge mid-points, quarter points
    o = np.array([[x, y, 0.0] for y in pts for x in pts], F32)
    d = np.tile(np.array([0, 0, 1], F32), (len(o), 1))
    s_ref, p_ref, t_ref, _ = oracle.OracleScene(scene).trace_rays(o, d)
    out, _ = run_harness(harness(variant), tmp_path, o, d, ["obj", obj])
    assert (s_ref != MISS).mean() > 0.9 and len(set(p_ref.tolist())) > 100
    assert np.array_equal(out["prim"], p_ref) and np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))
    # (2) rays that start on a surface (every scattered ray does) must not re-hit it below t = 1e-5
    name = "mushroom.obj"
    scene = Scene.empty()
    scene.add_object(scene.add_material(Materials.lambertian((0.5, 0.5, 0.5))), scene.add_mesh(load_obj(asset_path(name))))
    scene.environment = Environments.uniform((0.5, 0.5, 0.5))
    osc = oracle.OracleScene(scene)
    o, d = random_rays(30000, *scene_bounds(scene), seed=21)
    s0, _, t0, _ = osc.trace_rays(o, d)
    hit = s0 != MISS
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    o2 = (o[hit] + dn[hit] * t0[hit, None]).astype(F32)
    d2 = np.random.default_rng(5).normal(size=o2.shape).astype(F32)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o2, d2)
    out, _ = run_harness(harness(variant), tmp_path, o2, d2, ["obj", asset_path(name)])
    assert np.array_equal(out["surface"], s_ref) and np.array_equal(out["prim"], p_ref)
    assert np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))


# ---- scene-level culling of the reference (core/scene.rs:182-185 over core/bvh.rs:132-160, util/aabb.rs:86-148) ----
@pytest.mark.parametrize("case", ["two_coplanar_quads", "three_coplanar_quads", "five_coplanar_quads",
                                  "flat_mesh_and_sphere", "flat_pair_inside_larger_scene", "two_coplanar_walls_x"])
def test_scene_level_split_culling_matches_reference(oracle, harness, tmp_path, case):
    scene, args = build_case_scene(flat_split_cases(tmp_path)[case])
    o, d = flat_case_rays(scene, 3)
    osc = oracle.OracleScene(scene)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d, oracle.MODE_FAITHFUL)
    s_all, _, _, _ = osc.trace_rays(o, d, oracle.MODE_BRUTE)
    out, _ = run_harness(harness("default"), tmp_path, o, d, args)
    assert np.array_equal(out["surface"], s_ref)
    tri = p_ref != MISS
    assert np.array_equal(out["prim"][tri], p_ref[tri])
    assert np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))
    culled = (s_all != MISS) & (s_ref != s_all)
    if case == "flat_mesh_and_sphere":
        assert not culled.any() and (s_ref != MISS).mean() > 0.5
    else:
        # the case is not vacuous: the reference's flat Split drops hits that an un-culled closest hit finds
        assert culled.sum() > 100, int(culled.sum())


def test_scene_level_culling_beyond_the_mask_width(oracle, harness, tmp_path):
    # 40 surfaces (> the 32 visibility bits a ray carries): candidates walk their ancestor chain instead.
    # Coplanar tiles in rows of the plane y = 0 plus a second storey at y = 1, so flat and non-flat Splits mix.
    surfaces = []
    for k in range(40):
        p = str(tmp_path / f"tile_{k}.obj")
        x0, z0, y = 2.0 * (k % 8), 2.0 * ((k // 8) % 4), float(k // 32)
        quad_obj(p, [(x0, y, z0), (x0, y, z0 + 1), (x0 + 1, y, z0 + 1), (x0 + 1, y, z0)])
        surfaces.append(("obj", p))
    scene, args = build_case_scene(surfaces)
    o, d = flat_case_rays(scene, 9)
    osc = oracle.OracleScene(scene)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d, oracle.MODE_FAITHFUL)
    s_all, _, _, _ = osc.trace_rays(o, d, oracle.MODE_BRUTE)
    out, _ = run_harness(harness("default"), tmp_path, o, d, args)
    assert np.array_equal(out["surface"], s_ref) and np.array_equal(out["prim"], p_ref)
    assert np.array_equal(out["t"].view(np.uint32), t_ref.view(np.uint32))
    assert ((s_all != MISS) & (s_ref != s_all)).sum() > 100 and (s_ref != MISS).sum() > 100
