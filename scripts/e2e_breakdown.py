"""Where the e2e step time goes: H2D only, + extraction, + D2H of the results, + matching."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfmfromscratch_b200 import pipeline as PL
from sfmfromscratch_b200.extractor import extract_batch_device, make_params
from sfmfromscratch_b200.synth import synth_image
B, H, W, cap = 32, 1080, 1920, 2500
host = torch.from_numpy(np.stack([synth_image(H, W, s) for s in range(8)] * 4)).pin_memory()
dev = torch.device('cuda')
params, keep = make_params({}, pyramid=True)
imgs = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(2)]
i32 = dict(dtype=torch.int32, device=dev)
full = [{'x': torch.empty((B, cap), **i32), 'y': torch.empty((B, cap), **i32), 'count': torch.empty((B,), **i32),
         'desc': torch.empty((B, cap, 128), dtype=torch.float32, device=dev)} for _ in range(2)]
hout = [{k: torch.empty_like(v, device='cpu').pin_memory() for k, v in full[0].items()} for _ in range(2)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()
pairs = PL.consecutive_pairs(B); pp = torch.from_numpy(pairs).to(dev)
pipe = PL.FeaturePipeline({}, 0.8)
def step(k, do_extract, do_d2h, do_match, chunk=8):
    s = k & 1
    evs = []
    with torch.cuda.stream(s_in):
        for c0 in range(0, B, chunk):
            imgs[s][c0:c0 + chunk].copy_(host[c0:c0 + chunk], non_blocking=True)
            e = torch.cuda.Event(); e.record(s_in); evs.append(e)
    for c0, e in zip(range(0, B, chunk), evs):
        main.wait_event(e)
        if do_extract:
            extract_batch_device(imgs[s][c0:c0 + chunk], params, want_aux=False, check=False, out={k2: v[c0:c0 + chunk] for k2, v in full[s].items()})
        if do_d2h:
            d = torch.cuda.Event(); d.record(main); s_out.wait_event(d)
            with torch.cuda.stream(s_out):
                for k2 in ('x', 'y', 'desc', 'count'):
                    hout[s][k2][c0:c0 + chunk].copy_(full[s][k2][c0:c0 + chunk], non_blocking=True)
    if do_match:
        pipe.match(full[s]['desc'], full[s]['count'], pp, cap=cap, pairs_host=pairs)
    main.wait_stream(s_out)
def run(name, K=20, **kw):
    for k in range(3): step(k, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(main)
    for k in range(K): step(k, **kw)
    main.wait_stream(s_in)
    b.record(main)
    host_ms = (time.perf_counter() - t0) / K * 1e3
    torch.cuda.synchronize()
    print(f"{name:34s} {a.elapsed_time(b) / K:.3f} ms/step   (host enqueue {host_ms:.3f} ms/step)")
run("H2D only", do_extract=False, do_d2h=False, do_match=False)
run("H2D + extract", do_extract=True, do_d2h=False, do_match=False)
run("H2D + extract + D2H", do_extract=True, do_d2h=True, do_match=False)
run("H2D + extract + D2H + match", do_extract=True, do_d2h=True, do_match=True)
run("same, chunks of 4", do_extract=True, do_d2h=True, do_match=True, chunk=4)
run("same, chunks of 16", do_extract=True, do_d2h=True, do_match=True, chunk=16)
