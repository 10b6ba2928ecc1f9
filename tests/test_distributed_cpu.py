"""N>1 host logic on the CPU (gloo, world_size 2 and 3): tile sharding covers the grid exactly once in
order, and the rank-ordered variable-length gather reproduces the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "object-detection-yolov3_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import cases, tiling_np as tl
    from yolo3_b200 import gather_rows, shard_range, tile_count
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w, c, dt, tile, edge, nb, nc, minbox = cases.PIPE_CASES["e96"]
        img = cases.synthetic_image(h, w, c, dt, seed=100 + 3 + edge)
        fake = cases.FakeDetector(nb, nc, tile, seed=5 + edge)
        n_tiles = tile_count(h, w, tile, edge)
        first, count = shard_range(n_tiles, rank, world)
        # the oracle pipeline restricted to this rank's tiles: run everything, keep the rows that come
        # from tiles [first, first+count) - the reference's output is tile-major, so per-tile row counts
        # are obtained by running tile by tile
        tiles, xs, ys = tl.cut_tiles(img, tile, edge)
        rows = []
        for t in range(first, first + count):
            one = tl.tiled_inference_single_tile(fake, img, tile, minbox, edge, t)
            rows.append(one)
        local = torch.from_numpy(np.concatenate(rows) if rows else np.zeros((0, 6)))
        full = gather_rows(local)
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_single_process(golden, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, golden("tiled_pipeline.npz")["e96_pred"])


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "object-detection-yolov3_b200"))
    from yolo3_b200 import shard_range
    for n in (0, 1, 7, 2809, 3969):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for f, c in spans:
                assert c >= 0 and (c == 0 or f == pos)
                pos += c
