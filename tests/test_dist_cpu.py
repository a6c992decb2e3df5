"""CPU: the multi-GPU sharding / gather logic with world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracinggrff_b200 import dist as rdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, n_x, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx, rows = rdist.shard_rays(n_x, n_rows, world, rank)
        # stand-in for the renderer: value = f(frequency, flat ray index)
        mr = rdist.max_rows_per_rank(n_rows, world)
        local = torch.full((2, mr, n_x), -1.0, dtype=torch.float64)
        vals = torch.from_numpy(idx.astype(np.float64)).reshape(len(rows), n_x)
        local[0, :len(rows)] = vals
        local[1, :len(rows)] = 10.0 * vals
        full = rdist.gather_rows(local, n_rows)
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 8), (2, 7), (3, 10), (2, 150), (3, 200)])
def test_interleaved_rows_gather(world, n_rows):
    n_x = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, n_x, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_rows * n_x, dtype=np.float64).reshape(n_rows, n_x)
    for rank, full in outs:
        np.testing.assert_array_equal(full[0], expect)
        np.testing.assert_array_equal(full[1], 10.0 * expect)


def test_shard_covers_every_ray_once():
    for world in (1, 2, 4, 8):
        seen = np.concatenate([rdist.shard_rays(16, 37, world, r)[0] for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(16 * 37))
        sizes = [len(rdist.rows_of_rank(37, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) == rdist.max_rows_per_rank(37, world)
    # large images: rows go out in groups of 8 adjacent rows (the tile height), still every row exactly once,
    # shares within one group of each other
    for world, n_rows in ((2, 512), (8, 2048), (8, 4096), (3, 1000)):
        assert rdist.row_group(n_rows, world) == 8
        rows = [rdist.rows_of_rank(n_rows, world, r) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(n_rows))
        assert max(map(len, rows)) - min(map(len, rows)) <= 8
        assert np.array_equal(rows[1][:8], np.arange(8, 16))


def test_row_sharding_is_a_partition_for_any_size():
    """Every row exactly once, rank shares within one group of each other, groups of 8 only when every
    rank still gets at least 8 of them."""
    rng = np.random.default_rng(0)
    for _ in range(300):
        world = int(rng.integers(1, 9))
        n_rows = int(rng.integers(1, 5000))
        rows = [rdist.rows_of_rank(n_rows, world, r) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(n_rows))
        g = rdist.row_group(n_rows, world)
        assert g == (8 if n_rows >= 64 * world else 1)
        assert max(map(len, rows)) - min(map(len, rows)) <= g
        assert rdist.max_rows_per_rank(n_rows, world) == max(map(len, rows))
        for r in range(world):
            assert np.all(np.diff(rows[r]) > 0)
            idx, rr = rdist.shard_rays(7, n_rows, world, r)
            assert np.array_equal(rr, rows[r]) and idx.size == 7 * len(rows[r])


def test_c_abi_row_sharding_equals_the_python_partition():
    """rtgrff_shard_rows (host-only entry of the C ABI: what a non-Python host calls and what
    rtgrff_gather_image assumes) against dist.rows_of_rank for every world size the bench uses."""
    from raytracinggrff_b200 import dist as rdist
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 64, 129, 512, 2048):
            seen = []
            for r in range(world):
                rows, mr = rdist.c_shard_rows(n, world, r)
                assert np.array_equal(rows, rdist.rows_of_rank(n, world, r))
                assert mr == rdist.max_rows_per_rank(n, world)
                seen.append(rows)
            assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n))
