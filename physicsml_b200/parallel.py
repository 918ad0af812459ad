"""Data-parallel plumbing: one process per GPU, molecule batches sharded by rank, one gradient all-reduce per step
(the reference delegates this to Lightning DDP, lightning/model.py:159-163).  The ~1.4-6 M parameters of MACE fit one
flat bucket, so a single NCCL all-reduce over NVLink is launched after backward; it is latency- not bandwidth-bound."""
from __future__ import annotations

import torch
import torch.distributed as dist


def allreduce_gradients(params, world_size: int | None = None, group=None) -> None:
    """average .grad of `params` over the process group with ONE flat-bucket all-reduce (in place)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off : off + n].view_as(g))
        off += n


def shard_indices(n_items: int, rank: int, world_size: int) -> range:
    """DistributedSampler-style contiguous-stride shard of a dataset of n_items structures"""
    return range(rank, n_items, world_size)
