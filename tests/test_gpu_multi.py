"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process per GPU, sample-range sharding, then the root sums the
other rank's accumulation buffer through CUDA IPC + NVLink peer loads (vr_render_reduce_peers) and through the fused
reduce + resolve kernel; both must equal the single-GPU render of all samples."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, SPP = 160, 120, 16


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from voidray_b200 import scenes
    from voidray_b200.distributed import gather_accum_handles, reduce_accum_peers, shard_samples
    from voidray_b200.render import Context, RenderTarget
    from voidray_b200.scene import RenderSettings
    scene, st, _ = scenes.config1_mushroom(W, H, SPP)
    ctx = Context(rank)
    accel = scene.build_acceleration(ctx)
    off, cnt = shard_samples(SPP, world, rank)
    tgt = RenderTarget(accel, (W, H), RenderSettings(total_samples=SPP, max_bounces=8, sample_offset=off))
    tgt.accumulate(cnt)
    handles = gather_accum_handles(tgt, 0)
    dist.barrier()
    if rank == 0:
        fused = tgt.resolve_peers(handles, 1.0, 1.0, 1.0, 1)
    dist.barrier()
    reduce_accum_peers(tgt, handles, 0)
    if rank == 0:
        reduced = tgt.read()
        full = RenderTarget(accel, (W, H), RenderSettings(total_samples=SPP, max_bounces=8))
        full.accumulate(SPP)
        np.savez(out_path, reduced=reduced, fused=fused, full=full.read(), full_resolved=full.resolve(1.0, 1.0, 1.0, 1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_peer_reduce(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    d = np.load(out)
    assert np.abs(d["reduced"][..., :3] - d["full"][..., :3]).max() <= 2e-6
    assert np.all(d["reduced"][..., 3] == 2.0)
    assert np.allclose(d["fused"], d["full_resolved"], rtol=1e-4, atol=1e-5)


def test_inprocess_device_group_through_the_c_header(tmp_path):
    """vr_context_create_multi: one process, no torch — tests/c/multi_device.c renders on a device group and on one
    device through the C header alone and compares (accumulation buffer <= 2e-6, alpha exact, same segment count).
    Always run with two shards on device 0 (the sharding, replication and reduce logic on a one-GPU box); with two
    and with all devices over NVLink peer memory where the box has them."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "voidray_b200")
    exe = str(tmp_path / "multi_device")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "c", "multi_device.c"), "-L", libdir, "-lvoidray_cuda",
                        f"-Wl,-rpath,{libdir}", "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = torch.cuda.device_count()
    lists = ["0,0", "0,0,0"]
    if n >= 2:
        lists.append("0,1")
    if n > 2:
        lists.append(",".join(str(i) for i in range(n)))
    for ids in lists:
        r = subprocess.run([exe, os.path.join(root, "assets", "mushroom.obj"), ids], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0 and f"ok {len(ids.split(','))} devices" in r.stdout, (ids, r.stdout, r.stderr)


def test_inprocess_device_group_python_mirror(oracle):
    # Context.multi through the Python host mirror, against the oracle (fixed-seed per-pixel mean) on a two-surface scene
    from voidray_b200 import scenes
    from voidray_b200.render import Context, RenderTarget
    w, h, spp = 160, 90, 8
    scene, st, _ = scenes.config5_combined(w, h, spp)
    ref, c_ref = oracle.OracleScene(scene).render(w, h, st.render, spp)
    ctx = Context.multi([0, 0])
    tgt = RenderTarget(scene.build_acceleration(ctx), (w, h), st.render)
    tgt.accumulate(spp)
    img = tgt.read()
    assert np.abs(img[..., :3] - ref[..., :3]).max() <= 2e-4
    assert np.array_equal(img[..., 3], ref[..., 3])
    assert abs(tgt.stats().ray_segments - c_ref.segments) <= max(8, c_ref.segments // 20000)
