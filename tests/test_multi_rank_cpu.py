"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- the tile interleave that
svo_frame_desc.tile_rank / tile_world select (tile column tx -> rank (tx / 4) % world) partitions every pixel exactly once,
and the per-rank ray counts the bench all-reduces add up. No device compute."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cases, out_queue, run=4):
    sys.path.insert(0, str(ROOT / "sparse-voxel-octrees_b200"))
    import torch
    import pysvo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pysvo.frame_set_tile_run(run)        # stripe width in tile columns (svo_frame_set_tile_run; default 4)
    ok = True
    for (W, H, S) in cases:
        lay = pysvo.frame_layout(W, H, S)
        cover = np.zeros((H, W), np.int32)
        owned_pixels = 0
        mine = [t for t in range(lay.tiles) if pysvo.tile_owner(W, H, S, t, world) == rank]
        for t in mine:                                   # the tiles this rank's kernels take
            x0, y0, x1, y1 = pysvo.tile_rect(W, H, S, t)
            cover[y0:y1, x0:x1] += 1
            owned_pixels += (x1 - x0) * (y1 - y0)
        total = torch.from_numpy(cover.copy())
        dist.all_reduce(total)                           # every pixel must be owned exactly once overall
        px = torch.tensor([owned_pixels], dtype=torch.int64)
        dist.all_reduce(px)
        ok = ok and bool((total == 1).all()) and int(px.item()) == W * H and int(cover.max()) <= 1
        # load balance of the interleave (runs of 4 tile columns dealt round-robin): ranks differ by at most one run
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(mine)], dtype=torch.int64))
        rows = lay.tiles // lay.tile_cols
        ok = ok and max(int(c) for c in counts) - min(int(c) for c in counts) <= run * rows
    if rank == 0:
        out_queue.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,run", [(2, 4), (3, 4), (2, 16), (3, 15)])
def test_tile_interleave_partitions_the_frame(pysvo, world, run):
    """run = 4 is the rendering default; bench.py's host-visible leg at N > 1 deals wider stripes (15 / 16)."""
    cases = [(1280, 720, 16), (333, 77, 5), (64, 200, 7), (9, 9, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q, run)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_layout_matches_reference_formulas(pysvo):
    lay = pysvo.frame_layout(1280, 720, 16)
    assert (lay.n_strips, lay.strip_rows, lay.tiles_x, lay.tiles_y_full, lay.tiles_y_last) == (16, 45, 161, 7, 7)
    assert lay.corners == 18032 == pysvo.coarse_cells(1280, 720, 16)
    assert lay.tiles == 16 * 6 * 160
    lay = pysvo.frame_layout(3840, 2160, 16)
    assert lay.corners == 138528 and lay.tiles == 16 * 17 * 480
    for (W, H, S) in [(333, 77, 5), (64, 200, 7), (1, 1, 1), (1001, 13, 13)]:
        assert pysvo.frame_layout(W, H, S).corners == pysvo.coarse_cells(W, H, S)
