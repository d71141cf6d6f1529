"""Multi-GPU layout of a batch of MPC instances: contiguous shards, one all-gather of the solved policies per tick.

Instances are independent (own x0, reference, mode schedule, warm start: SURVEY.md section 8e), so the data path needs no
collective; the only exchange is the north star's all-gather of the feedback policies after each tick.
"""
from __future__ import annotations


def shard_range(batch_total: int, rank: int, world: int):
    """Contiguous block of instances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(batch_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_policies(local: dict, batch_total: int, world: int):
    """All-gathers every policy tensor (first dim = local instances) into [batch_total, ...] tensors on each rank."""
    import torch
    import torch.distributed as dist
    out = {}
    for name, t in local.items():
        sizes = [shard_range(batch_total, r, world) for r in range(world)]
        if all(b - a == sizes[0][1] - sizes[0][0] for a, b in sizes):
            full = torch.empty((batch_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(full, t.contiguous())
        else:
            parts = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for a, b in sizes]
            dist.all_gather(parts, t.contiguous())
            full = torch.cat(parts, dim=0)
        out[name] = full
    return out
