"""Multi-GPU plumbing: sample-range sharding + one sum-reduce of the accumulation buffers.

The reference has no multi-device path. Samples are i.i.d. and accumulation is a sum
(render/iterative.rs:45-51), so the path shards by sample range with no data-path exchange: rank r of N
renders global sample indices [offset_r, offset_r + count_r) of every pixel with the same
total_samples, and the per-rank accumulation buffers (already divided by total_samples) are summed once
by an NCCL reduce over NVLink (torch.distributed; gloo on CPU for the tests). Because the generator is
keyed by (pixel, global sample index), the union of the shards is exactly the 1-GPU sample set.
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_samples(total_samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous sample range of `rank`: (offset, count). The first total % world ranks get one extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(total_samples), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str = "nccl"):
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group(backend=backend)
    return dist


def reduce_accum(tensor, dst: int = 0, all_ranks: bool = False):
    """Sum the per-rank accumulation buffers: onto `dst` (ncclReduce) or everywhere (all-reduce)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensor
    if all_ranks:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor


def gather_accum_handles(target, dst: int = 0):
    """Exchange the CUDA IPC handles of every rank's accumulation buffer; returns the peers' handles (every rank
    but `dst`, in rank order) on `dst` and [] elsewhere. Done once per RenderTarget."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return []
    mine = torch.tensor(list(target.export_accum_handle()), dtype=torch.uint8,
                        device=f"cuda:{target.scene.ctx.device}" if dist.get_backend() == "nccl" else "cpu")
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    if dist.get_rank() != dst:
        return []
    return [bytes(t.cpu().tolist()) for r, t in enumerate(out) if r != dst]


def reduce_accum_peers(target, peer_handles, dst: int = 0):
    """One reduce step over peer memory: all ranks have finished accumulating (barrier), the root sums the peers'
    buffers into its own with NVLink peer loads, and nobody touches a buffer until it is done (barrier)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    dist.barrier()
    if dist.get_rank() == dst:
        target.reduce_peers(peer_handles)
    dist.barrier()
