"""N > 1 host logic on CPU: two `gloo` ranks shard the frame into interleaved row bands exactly as the CUDA backend does
(TileMap in csrc/rptr_cuda.cu), render their pixels with the oracle (global pixel ids in the RNG seed), and one
reduce(SUM) of the zero-initialised full-size accumulators on rank 0 reproduces the single-process image bit for bit
(disjoint support: sum == gather, exact in fp32)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, ROWS = 96, 70, 2, 8


def owned_rows(rank, world, rows=ROWS, height=H):
    return [y for y in range(height) if (y // rows) % world == rank]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from realtimepathtracingresearchframework_b200 import load_sky_fit, scenes
    s = scenes.random_triangles(3000, box=2.0, edge=0.4)
    s.camera = scenes.look_at_camera((0, 0, 7), (0, 0, 0))
    o = po.OracleScene(s)
    sp = load_sky_fit()
    img = np.zeros((H, W, 4), np.float32)
    rows = owned_rows(rank, world)
    # contiguous runs of owned rows = the bands of this rank
    start = prev = None
    for y in rows + [None]:
        if start is None:
            start = prev = y
        elif y is not None and y == prev + 1:
            prev = y
        else:
            o.render(W, H, s.camera, sp, spp=SPP, region=(0, start, W, prev + 1), out=img, n_threads=1)
            start = prev = y
    t = torch.from_numpy(img)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        full, _ = o.render(W, H, s.camera, sp, spp=SPP, n_threads=1)
        np.save(out_path, np.stack([t.numpy(), full]))
    dist.barrier()
    dist.destroy_process_group()


def test_band_partition_covers_every_row_once():
    for world in (1, 2, 3, 4, 8):
        rows = sorted(y for r in range(world) for y in owned_rows(r, world, height=1080))
        assert rows == list(range(1080))
        counts = [len(owned_rows(r, world, height=1080)) for r in range(world)]
        assert max(counts) - min(counts) <= ROWS


def test_two_rank_reduce_reproduces_single_process_image(oracle, tmp_path):
    out = str(tmp_path / "reduced.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    both = np.load(out)
    assert both[1][..., :3].max() > 0
    assert np.array_equal(both[0].view(np.uint32), both[1].view(np.uint32))
