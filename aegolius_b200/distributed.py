"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU for tests).

The path shards with NO data-path collective: every grid point depends only on its own coordinates and the (tiny,
replicated) program, so each rank evaluates one contiguous slab of ix planes — contiguous in SPOMSO's memory layout,
where x is the slowest axis (helper_functions.py:72-75). The only collective is the OPTIONAL assembly of the
full field on every rank (all_gather of equal-sized padded slabs); it is timed separately, never inside the throughput.
"""
from __future__ import annotations

import numpy as np

from .engine import slab_ranges


def rank_slab(n_planes: int, rank: int, world: int):
    """(x0, x1) of `rank`: np.array_split(range(n_planes), world)[rank] as a half-open range."""
    return slab_ranges(n_planes, world)[rank]


def padded_slab_points(spec, world: int) -> int:
    """Points of the largest slab (all_gather needs equal sizes; 1025 planes over 8 ranks = 129,128,...)."""
    per_plane = spec.res[1] * spec.res[2]
    return max(x1 - x0 for x0, x1 in slab_ranges(spec.res[0], world)) * per_plane


def gather_field(local, spec, group=None):
    """all_gather of the per-rank slabs into the full flat field (N,) on every rank.

    `local` is this rank's slab as a 1-D torch tensor (CUDA with NCCL, CPU with gloo). Slabs are padded to equal
    length for all_gather_into_tensor and the padding is dropped while assembling."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per_plane = spec.res[1] * spec.res[2]
    ranges = slab_ranges(spec.res[0], world)
    pad = padded_slab_points(spec, world)
    n_local = (ranges[rank][1] - ranges[rank][0]) * per_plane
    if local.numel() != n_local:
        raise ValueError(f"rank {rank}: slab has {local.numel()} points, expected {n_local}")
    send = local
    if n_local != pad:
        send = torch.zeros(pad, dtype=local.dtype, device=local.device)
        send[:n_local] = local
    recv = torch.empty(pad * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    out = torch.empty(spec.n_points, dtype=local.dtype, device=local.device)
    for r, (x0, x1) in enumerate(ranges):
        n_r = (x1 - x0) * per_plane
        out[x0 * per_plane:x0 * per_plane + n_r] = recv[r * pad:r * pad + n_r]
    return out


def evaluate_sharded(obj, spec, *, dtype="f32", grad=None, group=None, gather=False):
    """Evaluates this rank's slab on its GPU (device = LOCAL_RANK by default) and optionally gathers the field."""
    import os
    import torch
    import torch.distributed as dist
    from . import engine
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    device = int(os.environ.get("LOCAL_RANK", "0"))
    slab = rank_slab(spec.res[0], rank, world)
    res = engine.create_torch(obj, spec, dtype=dtype, grad=grad, device=device, slab=slab)
    if not gather or world == 1:
        return res
    field = res[0] if grad else res
    full = gather_field(field, spec, group)
    return (full, res[1]) if grad else full
