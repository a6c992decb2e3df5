"""Multi-GPU sharding of a map: one process per GPU (torch.distributed, NCCL over NVLink on the
GPUs, gloo in the CPU tests), replicated cubes, pixel rows dealt round-robin to the ranks, one
all-gather of the image slabs at the end.

This is the reference's only parallel strategy — data parallel over contiguous ray chunks with a
concatenate at the end (script/resample_with_ray_tracing.py:42-61, :333-352) — with two changes:
rows are interleaved instead of contiguous because disk-centre rays live much longer than limb
rays, and the exchange is a collective instead of pickles over pipes.  Rays never
interact, so there is no data-path collective during integration.
"""
from __future__ import annotations

import numpy as np


def rows_of_rank(n_rows: int, world_size: int, rank: int) -> np.ndarray:
    """Image rows owned by `rank`: rank, rank+W, rank+2W, ..."""
    return np.arange(rank, n_rows, world_size)


def max_rows_per_rank(n_rows: int, world_size: int) -> int:
    return (n_rows + world_size - 1) // world_size


def shard_rays(N_pix_x: int, N_pix_y: int, world_size: int, rank: int):
    """Flat ray indices (p = i*N_pix_x + j, script/resample_with_ray_tracing.py:470) of the rows of
    this rank, row-major within the shard."""
    rows = rows_of_rank(N_pix_y, world_size, rank)
    return (rows[:, None] * N_pix_x + np.arange(N_pix_x)[None, :]).ravel(), rows


def gather_rows(local, n_rows: int, group=None):
    """All-gather per-rank slabs ``local`` of shape (..., max_rows_per_rank, N_x) (rows beyond the
    rank's share are padding) into the full (..., n_rows, N_x) image on every rank.  `local` is a
    torch tensor on the device the process group's backend expects."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mr = max_rows_per_rank(n_rows, world)
    assert local.shape[-2] == mr, (local.shape, mr)
    local = local.contiguous()
    flat = local.reshape(-1)
    buf = torch.empty((world * flat.numel(),), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, flat, group=group)      # concatenation along dim 0 (gloo and NCCL)
    buf = buf.view((world,) + tuple(local.shape))
    full = torch.empty(tuple(local.shape[:-2]) + (n_rows, local.shape[-1]), dtype=local.dtype, device=local.device)
    for r in range(world):
        n_r = len(range(r, n_rows, world))
        full[..., r::world, :] = buf[r][..., :n_r, :]
    return full
