"""BASELINE.json configs[4] in full: all-pairs NN-ratio matching over 512 images x 8192 descriptors
(130 816 pairs), images' descriptor blocks sharded over the ranks, ONE all-gather, pairs dealt
block-cyclically.  Launch with torchrun for N > 1:

    python scripts/run_config5.py [--images 512] [--n 8192] [--chunk 256]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_config5.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmfromscratch_b200 import pipeline as PL  # noqa: E402
from sfmfromscratch_b200.matcher import match_batch_device, match_workspace  # noqa: E402


def synth_block(n_img, n, seed, dev):
    """RootSIFT-shaped descriptors generated on the device (numpy synthesis of 512 x 8192 rows takes minutes)."""
    g = torch.Generator(device=dev); g.manual_seed(1234)
    base = torch.rand((n, 128), generator=g, device=dev) ** 6
    g.manual_seed(seed)
    out = torch.empty((n_img, n, 128), device=dev)
    for i in range(n_img):
        h = torch.rand((n, 128), generator=g, device=dev) ** 6
        planted = torch.randperm(n, generator=g, device=dev)[: n // 2]
        src = torch.randperm(n, generator=g, device=dev)[: n // 2]
        h[planted] = base[src] + 0.03 * torch.rand((n // 2, 128), generator=g, device=dev) ** 2
        h = h / h.norm(dim=1, keepdim=True)
        out[i] = torch.sqrt(h)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=512)
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--chunk", type=int, default=256)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    s0, s1 = PL.shard_images(a.images, rank, world)
    per = (a.images + world - 1) // world
    mine = torch.zeros((per, a.n, 128), device=dev)
    mine[: s1 - s0] = synth_block(s1 - s0, a.n, 77 + rank, dev)
    counts = torch.zeros((per,), dtype=torch.int32, device=dev); counts[: s1 - s0] = a.n
    # one untimed chunk: kernel modules, the tensor-map entry point and the allocator's blocks are in place
    # before the clock starts (the timed region below still includes the all-gather and every pair)
    warm_pairs = torch.from_numpy(np.ascontiguousarray(PL.all_pairs(min(per, 24))[: a.chunk])).to(dev)
    if len(warm_pairs):
        wws = match_workspace(per, a.n, a.chunk, dev)
        match_batch_device(mine, counts.clamp(min=2), warm_pairs, 0.8, cap=a.n, ws=wws)
        del wws
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    desc_all, counts_all = PL.gather_descriptors(mine, counts)
    e1.record()
    pairs = PL.deal_pairs(PL.all_pairs(a.images), rank, world)
    total = torch.zeros((), dtype=torch.int64, device=dev)
    ws = match_workspace(desc_all.shape[0], a.n, a.chunk, dev)       # the sets are prepared by the first chunk only
    pairs_dev = torch.from_numpy(np.ascontiguousarray(pairs)).to(dev)
    for c0 in range(0, len(pairs), a.chunk):
        pc = pairs_dev[c0:c0 + a.chunk]
        m, c, cnt = match_batch_device(desc_all, counts_all, pc, 0.8, cap=a.n, ws=ws, prepared=c0 > 0)
        total += cnt.sum()          # stays on the device: no host sync per chunk
    e2.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], device=dev)
    tm = total.to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(tm)
    if rank == 0:
        npairs = a.images * (a.images - 1) // 2
        print(json.dumps({"config": f"all-pairs over {a.images} images x {a.n} descriptors", "n_gpus": world,
                          "image_pairs": npairs, "seconds": ms[0].item() / 1e3, "all_gather_ms": ms[1].item(),
                          "descriptor_pairs_per_s": npairs * a.n * a.n / (ms[0].item() / 1e3),
                          "algorithmic_tflops": 256.0 * npairs * a.n * a.n / (ms[0].item() / 1e3) / 1e12,
                          "matches": int(tm.item()), "wall_s": time.time() - t0}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
