"""Development aid: all-pairs chunks on one stream against two streams with private workspaces (does chunk k's
re-check overlap chunk k+1's tensor-core pass?)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sfmfromscratch_b200 import _native as N, pipeline as PL
from sfmfromscratch_b200.matcher import match_batch_device, match_workspace
N.load_library(); N.get_ctx(0)
dev = torch.device("cuda:0")
n_img, n, chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 96, 8192, 256
desc = bench.synth_descriptor_block(torch, n_img, n, 77, dev)
counts = torch.full((n_img,), n, dtype=torch.int32, device=dev)
pairs = torch.from_numpy(np.ascontiguousarray(PL.all_pairs(n_img))).to(dev)
P = len(pairs)

def run(nstreams):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    wss = [match_workspace(n_img, n, chunk, dev) for _ in range(nstreams)]
    totals = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(nstreams)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for i, c0 in enumerate(range(0, P, chunk)):
        k = i % nstreams
        with torch.cuda.stream(streams[k]):
            m, c, cnt = match_batch_device(desc, counts, pairs[c0:c0 + chunk], 0.8, cap=n, ws=wss[k], prepared=i >= nstreams)
            totals[k] += cnt.sum()
    for s in streams: main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), int(sum(t.item() for t in totals))

for ns in (1, 2, 1, 2, 3):
    ms, tot = run(ns)
    print(f"{ns} stream(s): {ms:.1f} ms, {P} pairs, {256.0 * P * n * n / (ms * 1e-3) / 1e12:.0f} TFLOP/s, matches {tot}")
