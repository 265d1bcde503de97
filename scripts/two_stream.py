"""Experiment: extraction steps alternated over two CUDA streams (memory-bound tail of one batch under the FMA-bound head of the next)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfmfromscratch_b200 import pipeline as PL
from sfmfromscratch_b200.synth import synth_image
B = 32
imgs = torch.from_numpy(np.stack([synth_image(1080, 1920, s) for s in range(8)] * 4)).cuda()
pairs = PL.consecutive_pairs(B)
pp = torch.from_numpy(pairs).cuda()
pipe = PL.FeaturePipeline({}, 0.8)
def step():
    out = pipe.extract(imgs)
    m = pipe.match(out['desc'], out['count'], pp, cap=2500, pairs_host=pairs)
    return out, m
def run(nstreams, K=40):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    main = torch.cuda.current_stream()
    for s in streams:
        with torch.cuda.stream(s):
            step(); step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for s in streams: s.wait_stream(main)
    for i in range(K):
        with torch.cuda.stream(streams[i % nstreams]):
            step()
    for s in streams: main.wait_stream(s)
    b.record(main)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K
for n in (1, 2, 3, 1, 2):
    print(n, "streams:", round(run(n), 4), "ms per step")
