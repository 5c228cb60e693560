"""-m gpu: behaviour of the C ABI / host mirror: call-order errors, accumulation semantics, batching
invariance, cancel, the progressive driver, sample-range sharding on one device."""
import os
import threading
import time

import numpy as np
import pytest

from voidray_b200 import _lib, scenes
from voidray_b200.distributed import shard_samples
from voidray_b200.render import (Context, PostProcessingData, PostProcessingPass, RenderAction, Renderer,
                                 RenderTarget, iterative_render)
from voidray_b200.scene import Environments, Materials, RenderSettings, Scene, Settings, Surfaces, Tonemap

from test_oracle_shading import sphere_scene
from util import F32

pytestmark = pytest.mark.gpu


def test_errors_are_statuses_not_crashes(ctx):
    lib = _lib.load()
    import ctypes as C
    # surfaces without objects: the reference would panic at the first hit (scene.rs:183-184)
    s = Scene.empty()
    s.add_analytic_surface(Surfaces.sphere((0, 0, 0), 1))
    with pytest.raises(_lib.VoidrayError) as e:
        s.build_acceleration(ctx)
    assert e.value.status == _lib.VR_ERR_INVALID and "objects[s]" in str(e.value)
    # render before commit, bad handles, bad settings
    h = C.c_void_p()
    _lib.check(lib.vr_scene_create(ctx.handle, C.byref(h)))
    out = C.c_uint32()
    assert lib.vr_scene_add_object(h, 3, 0, C.byref(out)) == _lib.VR_ERR_INVALID
    cs = _lib.RenderSettingsC(4, 4, 3.0, 0, 0, 0, 1, 0, 0)
    r = C.c_void_p()
    assert lib.vr_render_begin(h, 8, 8, C.byref(cs), C.byref(r)) == _lib.VR_ERR_INVALID
    assert b"commit" in lib.vr_last_error()
    _lib.check(lib.vr_scene_commit(h))
    assert lib.vr_render_begin(h, 0, 8, C.byref(cs), C.byref(r)) == _lib.VR_ERR_INVALID
    cs.total_samples = 0
    assert lib.vr_render_begin(h, 8, 8, C.byref(cs), C.byref(r)) == _lib.VR_ERR_INVALID
    idx = np.array([0, 1, 5], np.uint32)
    pos = np.zeros((3, 3), F32)
    assert lib.vr_scene_add_mesh(h, _lib.fptr(pos), None, None, 3, _lib.uptr(idx), 3, C.byref(out)) == _lib.VR_ERR_INVALID
    lib.vr_scene_destroy(h)
    with pytest.raises(_lib.VoidrayError):
        Context(99)


def test_accumulation_semantics_match_iterative_render(oracle, ctx):
    # two calls of 3 and 5 samples against total_samples = 8: alpha counts calls, values are sums / total
    scene, st, _ = scenes.config1_mushroom(96, 72, 8)
    rs = st.render
    osc = oracle.OracleScene(scene)
    ref, _ = osc.render(96, 72, rs, 3)
    ref, _ = osc.render(96, 72, rs, 5, accum=ref, sample_offset=3)
    accel = scene.build_acceleration(ctx)
    tgt = RenderTarget(accel, (96, 72), rs)
    iterative_render(tgt, accel, rs, 3)
    iterative_render(tgt, accel, rs, 5)
    img = tgt.read()
    assert np.all(img[..., 3] == 2.0)
    assert np.abs(img - ref).max() <= 2e-4
    st_ = tgt.stats()
    assert (st_.samples_done, st_.total_samples, st_.camera_samples) == (8, 8, 96 * 72 * 8)
    # one closest-hit launch per depth and batch; a call of >= 2 samples is cut into two batches (one per wavefront)
    assert st_.kernel_launches > 0 and st_.trace_launches == 2 * 2 * rs.max_bounces and st_.device_ms > 0 and st_.trace_ms > 0
    tgt.clear()
    assert not tgt.read().any() and tgt.stats().samples_done == 0


def test_wavefront_batching_is_invisible(ctx):
    # the number of paths in flight changes how a call is cut into wavefront batches, never the result
    scene, st, _ = scenes.config5_combined(128, 72, 12)
    accel = scene.build_acceleration(ctx)
    imgs = []
    for cap in (128 * 72, 128 * 72 * 5, 0):
        rs = RenderSettings(total_samples=12, max_bounces=8, max_paths_in_flight=cap)
        tgt = RenderTarget(accel, (128, 72), rs)
        tgt.accumulate(12)
        imgs.append(tgt.read())
    assert np.array_equal(imgs[0], imgs[1]) and np.array_equal(imgs[0], imgs[2])
    # and the render is deterministic run to run
    tgt = RenderTarget(accel, (128, 72), RenderSettings(total_samples=12, max_bounces=8))
    tgt.accumulate(12)
    assert np.array_equal(tgt.read(), imgs[0])


def test_two_wavefronts_are_invisible(ctx, monkeypatch):
    # consecutive batches alternate between two wavefronts on two streams; the per-pixel sums still run in batch order,
    # so the image is bit-identical to a single wavefront's (VOIDRAY_STREAMS=1, read when the render begins)
    scene, st, _ = scenes.config5_combined(160, 96, 24)
    accel = scene.build_acceleration(ctx)
    imgs, launches = [], []
    for streams in ("1", "2"):
        monkeypatch.setenv("VOIDRAY_STREAMS", streams)
        for cap in (160 * 96 * 4, 0):
            tgt = RenderTarget(accel, (160, 96), RenderSettings(total_samples=24, max_bounces=8, max_paths_in_flight=cap))
            tgt.accumulate(7)   # 7 + 17: odd batch counts, a last short batch
            tgt.accumulate(17)
            imgs.append(tgt.read())
            launches.append(tgt.stats().trace_launches)
            tgt.close()
    for img in imgs[1:]:
        assert np.array_equal(img, imgs[0])
    assert launches[0] != launches[2] or launches[1] != launches[3]  # the two settings really cut the work differently


def test_camera_ray_culling_is_invisible(ctx, monkeypatch):
    # k_raygen finishes the camera rays that miss the scene's bounds itself while the host sees that pay (a call that
    # culled less than 15 % of its camera rays switches it off for the calls that follow). Same image, same segment count
    # with it pinned off, pinned on (VOIDRAY_CAMERA_CULL, read when the render begins) and left to the host — on a scene
    # whose ground fills the frame (switched off after the first call) and on the mushroom alone (stays on)
    for recipe in (scenes.config5_combined, scenes.config1_mushroom):
        scene, st, _ = recipe(160, 96, 24)
        accel = scene.build_acceleration(ctx)
        imgs, segments = [], []
        for mode in ("0", "1", None):
            if mode is None:
                monkeypatch.delenv("VOIDRAY_CAMERA_CULL", raising=False)
            else:
                monkeypatch.setenv("VOIDRAY_CAMERA_CULL", mode)
            tgt = RenderTarget(accel, (160, 96), RenderSettings(total_samples=24, max_bounces=8))
            tgt.accumulate(1)
            tgt.accumulate(23)
            imgs.append(tgt.read())
            segments.append(tgt.stats().ray_segments)
            tgt.close()
        assert np.array_equal(imgs[0], imgs[1]) and np.array_equal(imgs[0], imgs[2])
        assert segments[0] == segments[1] == segments[2]  # culled camera rays are segments all the same


def test_sample_range_sharding_on_one_device(ctx):
    # N renders owning disjoint sample ranges, summed, equal one render of all samples (up to f32 summation order)
    w, h, spp = 128, 72, 16
    scene, st, _ = scenes.config1_mushroom(w, h, spp)
    accel = scene.build_acceleration(ctx)
    full = RenderTarget(accel, (w, h), RenderSettings(total_samples=spp, max_bounces=8))
    full.accumulate(spp)
    want = full.read()
    for world in (2, 4):
        total = np.zeros_like(want)
        for rank in range(world):
            off, cnt = shard_samples(spp, world, rank)
            t = RenderTarget(accel, (w, h), RenderSettings(total_samples=spp, max_bounces=8, sample_offset=off))
            t.accumulate(cnt)
            total += t.read()
        assert np.abs(total[..., :3] - want[..., :3]).max() <= 2e-6
        assert np.all(total[..., 3] == world)
    # the accumulation buffer is reachable as a torch tensor for NCCL reduces
    import torch
    ten = full.as_torch()
    assert ten.is_cuda and ten.numel() == w * h * 4
    assert np.array_equal(ten.cpu().numpy().reshape(h, w, 4), want)


def test_cancel_leaves_whole_samples(ctx):
    scene, st, _ = scenes.config5_combined(640, 360, 4096)
    rs = RenderSettings(total_samples=4096, max_bounces=8, max_paths_in_flight=640 * 360)
    accel = scene.build_acceleration(ctx)
    tgt = RenderTarget(accel, (640, 360), rs)
    timer = threading.Timer(0.05, tgt.cancel)
    timer.start()
    with pytest.raises(_lib.RenderCancelled):
        tgt.accumulate(4096)
    timer.join()
    done = tgt.stats().samples_done
    assert 0 < done < 4096
    img = tgt.read()
    assert np.all(img[..., 3] == 1.0) and np.all(np.isfinite(img))
    # what is in the buffer is exactly `done` samples
    ref = RenderTarget(accel, (640, 360), rs)
    ref.accumulate(done)
    assert np.array_equal(ref.read(), img)
    tgt.accumulate(1)     # usable again after a cancel
    assert tgt.stats().samples_done == done + 1


def test_renderer_one_shot_driver(ctx):
    # RenderThread::one_shot (renderer.rs:35-125): 1-spp probe, batches, stats, post-process scale = total/done
    scene, st, dims = scenes.config1_mushroom(160, 120, 40)
    st.render.update_frequency = 0.005
    r = Renderer(scene, st, dims, ctx)
    with pytest.raises(RuntimeError):
        r.execute(RenderAction.Cancel)          # nothing running: the reference panics (renderer.rs:229-231)
    r.execute(RenderAction.Render)
    r.join()
    assert r.samples() == (40, 40) and not r.currently_rendering()
    assert r.elapsed_time() > 0 and r.remaining_time() is None
    img = r.post_process()
    assert img.shape == (120, 160, 4) and np.all(np.isfinite(img)) and img[..., :3].max() > 0.1
    want = PostProcessingPass().render(r.target, PostProcessingData(1.0, 1.0, 1.0, int(Tonemap.ACES)))
    assert np.array_equal(img, want)


def test_native_obj_loader_and_integer_textures(oracle, ctx):
    # vr_scene_add_mesh_from_obj_file (obj-rs semantics in C++) against the Python loader feeding the oracle, and
    # vr_scene_add_texture_rgb8 (`to_rgb32f` = byte / 255) against the f32 entry point
    import ctypes as C
    from PIL import Image
    from voidray_b200.assets import asset_path, load_obj
    from voidray_b200.scene import Camera
    from util import random_rays, scene_bounds
    lib = _lib.load()
    for name in ("cube.obj", "mushroom.obj", "fancy_monkey.obj"):
        scene = Scene.empty()
        sf = scene.add_mesh_from_file(asset_path(name))
        scene.add_object(scene.add_material(Materials.lambertian((0.5, 0.5, 0.5))), sf)
        accel = scene.build_acceleration(ctx)
        mesh = load_obj(asset_path(name))
        assert accel.info()["n_triangles"] == mesh.n_triangles
        osc = oracle.OracleScene(scene)
        assert np.array_equal(accel.tie_ranks(0), osc.global_tie_rank(0))
        o, d = random_rays(50000, mesh.positions.min(0), mesh.positions.max(0), seed=2)
        s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
        s, p, t = accel.trace_rays(o, d)
        assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    h = C.c_void_p()
    _lib.check(lib.vr_scene_create(ctx.handle, C.byref(h)))
    out = C.c_uint32()
    assert lib.vr_scene_add_mesh_from_obj_file(h, b"/nonexistent.obj", C.byref(out), None, None) == _lib.VR_ERR_INVALID
    img8 = np.ascontiguousarray(np.asarray(Image.open(asset_path("uv_test.png")).convert("RGBA")))
    _lib.check(lib.vr_scene_add_texture_rgb8(h, img8.ctypes.data_as(C.POINTER(C.c_uint8)), img8.shape[1], img8.shape[0],
                                             4, 1, C.byref(out)))
    f32 = np.ascontiguousarray(img8[..., :3].astype(F32) / F32(255.0))
    _lib.check(lib.vr_scene_add_texture_rgb32f(h, _lib.fptr(f32), f32.shape[1], f32.shape[0], 1, C.byref(out)))
    _lib.check(lib.vr_scene_commit(h))
    uv = np.random.default_rng(3).uniform(-1, 2, (5000, 2)).astype(F32)
    a = np.empty((5000, 3), F32)
    b = np.empty((5000, 3), F32)
    _lib.check(lib.vr_debug_texture_sample(h, 0, 5000, _lib.fptr(uv), _lib.fptr(a)))
    _lib.check(lib.vr_debug_texture_sample(h, 1, 5000, _lib.fptr(uv), _lib.fptr(b)))
    assert np.array_equal(a, b)
    lib.vr_scene_destroy(h)


def test_image_file_entry_points(ctx, tmp_path):
    # vr_scene_add_image_texture_file / vr_scene_set_environment_hdri_file (Scene::add_image_texture,
    # Environments::hdri: decode inside the library) against the array entry points fed by the same decoder
    import ctypes as C
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    from voidray_b200.assets import asset_path, load_image_native, synth_hdri
    lib = _lib.load()
    hdri_path = str(tmp_path / "env.exr")
    assert cv2.imwrite(hdri_path, synth_hdri("indoor", 256, 128)[:, :, ::-1].astype(F32))
    h = C.c_void_p()
    _lib.check(lib.vr_scene_create(ctx.handle, C.byref(h)))
    out = C.c_uint32()
    names = ("mushroom_albedo.jpg", "wood_normal.tif", "uv_test.png")
    for i, name in enumerate(names):
        _lib.check(lib.vr_scene_add_image_texture_file(h, os.fsencode(asset_path(name)), 1, C.byref(out)))
        assert out.value == 2 * i
        img = load_image_native(asset_path(name))
        _lib.check(lib.vr_scene_add_texture_rgb32f(h, _lib.fptr(img), img.shape[1], img.shape[0], 1, C.byref(out)))
    assert lib.vr_scene_add_image_texture_file(h, b"/nonexistent.png", 1, C.byref(out)) == _lib.VR_ERR_INVALID
    assert b"cannot open" in lib.vr_last_error()
    assert lib.vr_scene_set_environment_hdri_file(h, os.fsencode(asset_path("cube.obj"))) == _lib.VR_ERR_INVALID
    _lib.check(lib.vr_scene_set_environment_hdri_file(h, os.fsencode(hdri_path)))
    _lib.check(lib.vr_scene_commit(h))
    uv = np.random.default_rng(5).uniform(-1, 2, (4000, 2)).astype(F32)
    for i in range(len(names)):
        a, b = np.empty((4000, 3), F32), np.empty((4000, 3), F32)
        _lib.check(lib.vr_debug_texture_sample(h, 2 * i, 4000, _lib.fptr(uv), _lib.fptr(a)))
        _lib.check(lib.vr_debug_texture_sample(h, 2 * i + 1, 4000, _lib.fptr(uv), _lib.fptr(b)))
        assert np.array_equal(a, b) and a.max() > 0.1
    d = np.random.default_rng(6).normal(size=(4000, 3)).astype(F32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(F32)
    e_file = np.empty((4000, 3), F32)
    _lib.check(lib.vr_debug_environment_sample(h, 4000, _lib.fptr(d), _lib.fptr(e_file)))
    env = load_image_native(hdri_path)
    assert np.array_equal(env, synth_hdri("indoor", 256, 128).astype(F32))
    _lib.check(lib.vr_scene_set_environment_hdri_rgb32f(h, _lib.fptr(env), 256, 128))
    _lib.check(lib.vr_scene_commit(h))
    e_arr = np.empty((4000, 3), F32)
    _lib.check(lib.vr_debug_environment_sample(h, 4000, _lib.fptr(d), _lib.fptr(e_arr)))
    assert np.array_equal(e_file, e_arr) and e_file.max() > 1.0
    lib.vr_scene_destroy(h)


def test_peer_memory_reduce_and_fused_resolve(oracle, ctx):
    # the multi-GPU reduce kernels on peers living in this process: three renders own disjoint sample ranges; the
    # root sums the other two in list order (bit-exact against numpy in the same order) and the fused
    # reduce + resolve equals resolve(sum)
    w, h, spp = 160, 90, 12
    scene, st, _ = scenes.config1_mushroom(w, h, spp)
    accel = scene.build_acceleration(ctx)
    parts = []
    for rank in range(3):
        off, cnt = shard_samples(spp, 3, rank)
        t = RenderTarget(accel, (w, h), RenderSettings(total_samples=spp, max_bounces=8, sample_offset=off))
        t.accumulate(cnt)
        parts.append(t)
    bufs = [t.read() for t in parts]
    want = (bufs[0] + bufs[1]) + bufs[2]
    for mode in (0, 1, 3):
        fused = parts[0].resolve_peer_targets(parts[1:], 1.0, 2.2, 0.5, mode)
        ref = oracle.resolve(want, 1.0, 2.2, 0.5, mode)
        both_nan = np.isnan(fused) & np.isnan(ref)
        assert np.all(np.isclose(fused, ref, rtol=1e-5, atol=1e-6) | both_nan)
    assert np.array_equal(parts[0].read(), bufs[0])            # the fused kernel does not touch the accumulation buffer
    parts[0].reduce_peer_targets(parts[1:])
    assert np.array_equal(parts[0].read(), want)
    full = RenderTarget(accel, (w, h), RenderSettings(total_samples=spp, max_bounces=8))
    full.accumulate(spp)
    assert np.abs(full.read()[..., :3] - want[..., :3]).max() <= 2e-6
    handle = parts[1].export_accum_handle()
    assert len(handle) == 64 and any(handle)
