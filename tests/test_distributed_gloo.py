"""world_size-2 gloo test of the N>1 host logic: sample-range sharding + one sum-reduce reproduces the
single-process accumulation buffer. The per-rank renderer here is the CPU oracle (no GPU in this tier);
on the GPU box tests/test_gpu_sharding.py checks the same property through the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voidray_b200.distributed import reduce_accum, shard_samples
from voidray_b200.scene import RenderSettings

W, H, SPP = 24, 18, 7


def _scene():
    from test_oracle_shading import sphere_scene
    from voidray_b200.scene import Materials
    return sphere_scene(Materials.lambertian((0.6, 0.4, 0.2)), env=(0.7, 0.8, 0.9))


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    osc = O.OracleScene(_scene())
    rs = RenderSettings(total_samples=SPP, max_bounces=6)
    offset, count = shard_samples(SPP, world, rank)
    acc, _ = osc.render(W, H, rs, count, sample_offset=offset, n_threads=1)
    t = torch.from_numpy(acc.reshape(-1))
    reduce_accum(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy().reshape(H, W, 4))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sample_sharding_matches_single_rank(tmp_path, oracle):
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sharded = np.load(out)
    rs = RenderSettings(total_samples=SPP, max_bounces=6)
    single, _ = oracle.OracleScene(_scene()).render(W, H, rs, SPP, n_threads=1)
    # same sample set; only the f32 summation order differs ((a+b+c+d)+(e+f+g) vs a+...+g)
    assert np.allclose(sharded[..., :3], single[..., :3], rtol=0, atol=2e-6)
    assert np.all(sharded[..., 3] == 2.0)       # one `alpha += 1` per rank (iterative.rs:51)
    assert np.abs(sharded[..., :3] - single[..., :3]).max() < 2e-6
