"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard paths and row bands with
footile_b200.sharding and gather the bands; the rendered content comes from the oracle (the CUDA
path cannot run here), so what is tested is the partitioning and the gather, not the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from footile_b200.sharding import band_rows, shard_range
    for n in (0, 1, 7, 64, 100000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert spans[-1][0] + spans[-1][1] == n
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    for h in (1, 24, 100, 4096, 32768):
        for world in (1, 2, 4, 8):
            b = [band_rows(h, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == h
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert all(x[0] % 8 == 0 or x[0] == h for x in b)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from footile_b200 import scenes
    from footile_b200.sharding import band_rows, gather_bands, max_over_ranks, shard_range
    # (1) one raster split into row bands, gathered: equals the unsplit raster
    size = 96
    ops = scenes.random_polygons(3, 12, vertices=9, size=size, extent=48)
    full = oracle.Plotter(size, size, oracle.MATTE8, vid_cap=1 << 30, orderfree=True).fill(1, ops, (255,)).raster()
    b, e = band_rows(size, rank, world)
    got = gather_bands(torch.from_numpy(full[b:e].copy()), size, size)
    assert np.array_equal(got.numpy(), full)
    # (2) independent paths: every path drawn exactly once, by the rank that owns it
    n = 11
    first, count = shard_range(n, rank, world)
    mine, _, _ = scenes.random_curve_paths(first, count, segments=4, size=64)
    allp, offs, _ = scenes.random_curve_paths(0, n, segments=4, size=64)
    assert mine.tobytes() == allp[int(offs[first]): int(offs[first + count])].tobytes()
    owned = torch.zeros(n, dtype=torch.int32)
    owned[first: first + count] = 1
    dist.all_reduce(owned)
    assert owned.tolist() == [1] * n
    # (3) timing is the max over ranks
    assert max_over_ranks(1.0 + rank) == float(world)
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").close()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
