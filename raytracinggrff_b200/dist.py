"""Multi-GPU sharding of a map: one process per GPU (torch.distributed, NCCL over NVLink on the
GPUs, gloo in the CPU tests), replicated cubes, pixel rows dealt round-robin to the ranks, one
all-gather of the image slabs at the end.

This is the reference's only parallel strategy — data parallel over contiguous ray chunks with a
concatenate at the end (script/resample_with_ray_tracing.py:42-61, :333-352) — with two changes:
rows are interleaved instead of contiguous because disk-centre rays live much longer than limb
rays — in groups of 8 adjacent rows, the height of the pixel tile a warp walks
(RaySession.render_map), so that a warp's 32 rays stay neighbours in the real image — and the
exchange is a collective instead of pickles over pipes.  Rays never
interact, so there is no data-path collective during integration.
"""
from __future__ import annotations

import numpy as np


ROW_GROUP = 8      # = the tile height of RaySession.render_map(tile=(4, 8))


def row_group(n_rows: int, world_size: int) -> int:
    """Rows dealt together: 8 when every rank still gets at least 8 groups (load balance), else 1."""
    return ROW_GROUP if n_rows >= 8 * ROW_GROUP * world_size else 1


def rows_of_rank(n_rows: int, world_size: int, rank: int) -> np.ndarray:
    """Image rows owned by `rank`: groups of row_group() adjacent rows dealt round-robin
    (group g -> rank g mod W); with groups of one row that is rank, rank+W, rank+2W, ..."""
    g = np.arange(n_rows) // row_group(n_rows, world_size)
    return np.flatnonzero(g % world_size == rank)


def max_rows_per_rank(n_rows: int, world_size: int) -> int:
    return max(len(rows_of_rank(n_rows, world_size, r)) for r in range(world_size))


def shard_rays(N_pix_x: int, N_pix_y: int, world_size: int, rank: int):
    """Flat ray indices (p = i*N_pix_x + j, script/resample_with_ray_tracing.py:470) of the rows of
    this rank, row-major within the shard."""
    rows = rows_of_rank(N_pix_y, world_size, rank)
    return (rows[:, None] * N_pix_x + np.arange(N_pix_x)[None, :]).ravel(), rows


def c_shard_rows(n_rows: int, world_size: int, rank: int):
    """The same partition from the C ABI (rtgrff_shard_rows), which is what a non-Python host calls and what
    rtgrff_gather_image assumes: (rows, max_rows_per_rank)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    rows = np.empty(max(n_rows, 1), dtype=np.int32)
    n_local, max_rows = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(lib.rtgrff_shard_rows(int(n_rows), int(world_size), int(rank), _lib.ptr(rows, ctypes.c_int32),
                                     ctypes.byref(n_local), ctypes.byref(max_rows)))
    return rows[:n_local.value].astype(np.int64), int(max_rows.value)


def init_comm(session, group=None):
    """NCCL communicator of the library over all ranks of the torch.distributed group (any backend):
    rank 0 draws the unique id (rtgrff_comm_unique_id), torch.distributed carries its 128 bytes to
    the others, every rank joins (rtgrff_comm_init_rank).  The image then travels through
    ``RaySession.gather_image`` (rtgrff_gather_image), not through torch."""
    import ctypes
    import torch.distributed as dist
    from . import _lib
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    box = [None]
    if rank == 0 and world > 1:
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().rtgrff_comm_unique_id(buf))
        box[0] = buf.raw
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    session.comm_init(world, rank, box[0])
    return world, rank


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
        rows = torch.from_numpy(rows_of_rank(n_rows, world, r)).to(local.device)
        full[..., rows, :] = buf[r][..., :len(rows), :]
    return full
