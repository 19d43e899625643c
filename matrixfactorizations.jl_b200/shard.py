"""Multi-GPU sharding of batched work by contiguous matrix-index ranges.

Batched QL problems are independent, so there is NO data-path collective: rank g of G owns
matrices [g*N/G, (g+1)*N/G) (+ remainder spread over the first ranks), factors them on its own GPU
and the results stay sharded; an optional final gather (torch.distributed all_gather of the
per-rank slabs) is the only communication.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) of `total` matrices owned by `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_shards(local, total: int, group=None):
    """Final gather: concatenate every rank's slab (first dim = matrix index) on all ranks.
    Works on gloo (CPU tensors) and nccl (CUDA tensors); pads to the largest shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
