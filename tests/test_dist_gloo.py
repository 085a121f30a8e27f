"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding and the single
compact gather of the detect stage (refinedet/pytorch_b200/dist.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from refinedet.pytorch_b200 import dist as rdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_detections(rank, b_loc, C, seed):
    g = torch.Generator().manual_seed(seed + rank)
    counts = torch.randint(0, 6, (b_loc, C), generator=g, dtype=torch.int32)
    counts[:, 0] = 0
    total = int(counts.sum())
    rows = torch.rand(total, 5, generator=g) + rank           # rank-tagged so mix-ups show
    return counts, rows


def _worker(rank, world, port, num_images, C, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = rdist.shard_range(num_images, rank, world)
        counts, rows = _fake_detections(rank, hi - lo, C, 99)
        counts_all, rows_all = rdist.gather_packed(counts, rows)
        assert len(counts_all) == world and len(rows_all) == world
        for r in range(world):
            rlo, rhi = rdist.shard_range(num_images, r, world)
            ec, er = _fake_detections(r, rhi - rlo, C, 99)
            assert torch.equal(counts_all[r], ec), (rank, r)
            assert torch.equal(rows_all[r], er), (rank, r)
        np.save(os.path.join(out_dir, 'ok_%d.npy' % rank), np.array([sum(int(c.sum()) for c in counts_all)]))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_the_batch():
    for n in (1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [rdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_packed_world2_gloo(tmp_path):
    world, num_images, C = 2, 5, 4          # ragged: ranks own 3 and 2 images
    mp.spawn(_worker, args=(world, _free_port(), num_images, C, str(tmp_path)), nprocs=world, join=True)
    totals = [int(np.load(os.path.join(str(tmp_path), 'ok_%d.npy' % r))[0]) for r in range(world)]
    assert totals[0] == totals[1] > 0


def test_gather_packed_without_process_group_is_identity():
    counts, rows = _fake_detections(0, 3, 4, 1)
    c, r = rdist.gather_packed(counts, rows)
    assert torch.equal(c[0], counts) and torch.equal(r[0], rows)
