"""Where a multi-GPU step's time goes (development aid; launch with torchrun): extraction alone, the
pair plan's exchange alone, matching alone, and the whole step, each over 100 iterations."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfmfromscratch_b200 import pipeline as PL
from sfmfromscratch_b200.matcher import match_batch_device
from sfmfromscratch_b200.synth import synth_image

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
B = 32
imgs = torch.from_numpy(np.stack([synth_image(1080, 1920, 1000 * rank + s) for s in range(8)] * 4)).to(dev)
pipe = PL.FeaturePipeline({}, 0.8, rank=rank, world=world)
pairs = PL.consecutive_pairs(world * B)
plan = pipe.pair_plan(pairs, B)
out = pipe.extract(imgs)
tab = PL.exchange_for(plan, out['desc'], out['count'], None)


def timed(fn, n=100):
    for _ in range(5):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def full():
    o = pipe.extract(imgs)
    pipe.match_plan(plan, o['desc'], o['count'], cap=2500)


def full_allgather():
    o = pipe.extract(imgs)
    d, c = pipe.exchange(o['desc'], o['count'])
    pipe.match(d, c, torch.from_numpy(np.ascontiguousarray(PL.deal_pairs(pairs, rank, world, block=4))).to(dev), cap=2500)


res = {
    "extract": timed(lambda: pipe.extract(imgs)),
    "exchange_plan": timed(lambda: PL.exchange_for(plan, out['desc'], out['count'], None)),
    "exchange_all": timed(lambda: pipe.exchange(out['desc'], out['count'])),
    "match_plan_table": timed(lambda: match_batch_device(tab[0], tab[1], plan.pairs_dev(dev), 0.8, cap=2500)),
    "step_plan": timed(full),
    "step_allgather": timed(full_allgather),
}
if rank == 0:
    print(world, {k: round(v, 4) for k, v in res.items()}, "policy", plan.policy, "K", plan.K)
if world > 1:
    dist.destroy_process_group()
