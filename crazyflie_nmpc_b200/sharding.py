"""Batch sharding across GPUs (one process per GPU, torch.distributed).

Instances are independent (the reference has no batch notion at all): ranks own contiguous
slices of the batch, there is no exchange step inside the solve, and the only collective is the optional
all-gather of the solved first controls u0 (4 doubles per instance) over NCCL / NVLink.
"""
import numpy as np


def shard_bounds(B, world, rank):
    """Contiguous slice [lo, hi) of a batch of B instances owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    q, r = divmod(B, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_workload(w, world, rank):
    """Slice every batch array of a workload dict (crazyflie_nmpc_b200.workloads)."""
    B = w["x0"].shape[0]
    lo, hi = shard_bounds(B, world, rank)
    return {k: np.ascontiguousarray(v[lo:hi]) if hasattr(v, "shape") and v.shape[:1] == (B,) else v for k, v in w.items()}


def gather_u0(u0_local, B, group=None):
    """All-gather the per-rank u0 [b_rank, 4] tensors into [B, 4] on every rank (ragged shards are padded
    to the largest shard for the collective and trimmed afterwards)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bmax = -(-B // world)
    pad = torch.zeros(bmax, 4, dtype=u0_local.dtype, device=u0_local.device)
    pad[: u0_local.shape[0]] = u0_local
    out = torch.empty(world * bmax, 4, dtype=u0_local.dtype, device=u0_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(B, world, r)
        parts.append(out[r * bmax: r * bmax + (hi - lo)])
    return torch.cat(parts, dim=0)
