#!/usr/bin/env python
"""bench.py -- throughput of the SfmFromScratch feature hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the hot path over one batch per GPU: SIFT extraction
(ScaleRotInvSIFT, reference defaults) of 32 synthetic 1920x1080 images
(BASELINE.json configs[1] image shape, configs[3]'s per-GPU share at 8 GPUs),
the all-gather of the descriptor blocks when N > 1, and NN-ratio matching of
this rank's share of the consecutive image pairs (what Runner.py:183-191
matches).  `value` is input Mpixel/s of the whole job with the images resident
in HBM; `e2e` repeats the step through host buffers (pinned H2D of the images,
D2H of keypoints, descriptors and matches inside the timed region).  The
matcher is also measured alone on configs[4]-shaped pairs (8192 x 8192
descriptors) and reported under "match" with its tensor-pipe roofline.

The reference arm (--impl reference) times the CPU oracle port of the same
path (the reference is pure Python and cannot travel to the GPU box; see
DESIGN.md) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W, BATCH = 1080, 1920, 32
N_DISTINCT = 8                  # distinct synthetic images, tiled to the batch
MATCH_SETS, MATCH_N = 12, 8192  # matcher-only leg: all 66 pairs of 12 sets of 8192 descriptors
RATIO = 0.8
METRIC = "sift_extract_plus_nn_ratio_match_input_mpixel_per_s"
WORKLOAD = (f"ScaleRotInvSIFT(defaults: 4 levels, k=2500) on {BATCH} synthetic {IMG_W}x{IMG_H} f32 images per GPU "
            f"(configs[1] image, configs[3] per-GPU share) + NN-ratio matching (thr {RATIO}) of consecutive image pairs")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def level_pixels(h, w, levels=4, f=2):
    tot = 0
    for _ in range(levels):
        tot += h * w
        h, w = int(h / f), int(w / f)
    return tot


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and "None" not in nm and "All" not in nm:
                        self.reasons.add(nm.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------- CPU arm (oracle port)

def _cpu_init():
    from oracle import oracle as O
    O.lib()


def _cpu_extract(img):
    from oracle import oracle as O
    e = O.ScaleRotInvSIFT(img, {})
    return e.extract_descriptors()


def _cpu_match(args):
    from oracle import oracle as O
    f1, f2 = args
    return len(O.NNRatioFeatureMatcher(RATIO).match_features_ratio_test(f1, f2)[0])


def cpu_images(n_images):
    from sfmfromscratch_b200.synth import synth_image
    return [synth_image(IMG_H, IMG_W, s) for s in range(n_images)]


def cpu_sample(images, pool=None):
    """The oracle port on `images` (1080p) + their consecutive-pair matches, in
    this process or over `pool`; returns (Mpixel/s, seconds).  Input synthesis
    and worker start-up are outside the timed region."""
    n = len(images)
    t0 = time.time()
    if pool is None:
        feats = [_cpu_extract(im) for im in images]
        for i in range(n - 1):
            _cpu_match((feats[i], feats[i + 1]))
    else:
        feats = pool.map(_cpu_extract, images, chunksize=1)
        pool.map(_cpu_match, [(feats[i], feats[i + 1]) for i in range(n - 1)], chunksize=1)
    dt = time.time() - t0
    return n * IMG_H * IMG_W / dt / 1e6, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    n = max(cores, 2)
    images = cpu_images(n)
    steps = max(1, min(args.steps, 3))          # bounded: each step is a sample of the workload
    vals, times = [], []
    with mp.get_context("spawn").Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_extract, images[:cores], chunksize=1)     # warm-up: imports, page-in
        for _ in range(steps):
            v, dt = cpu_sample(images, pool)
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    sample = (f"each step = {n} of the workload's {BATCH} 1080p images + their {n - 1} consecutive-pair matches, "
              f"oracle port (C + numpy restatement of the reference; the reference itself is pure Python and is not on this box), "
              f"{cores} worker processes; {steps} steps timed")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": float(np.mean(times)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- SURVEY 8f rows 2-3

def geometry_leg(torch, N, local, cpu=True):
    """Two-view RANSAC (Runner.py:170,351: 5 967 hypotheses per pair) over 16 synthetic match sets of
    600..2500 correspondences, and the association scans, each timed (a) on the device with inputs
    and 8-subsets resident, (b) through the reference-facing call with numpy in/out, which includes
    the host replay of numpy's MT19937 stream that draws the subsets."""
    from sfmfromscratch_b200 import association as AS
    from sfmfromscratch_b200 import geometry as GE
    from sfmfromscratch_b200.synth import two_view_correspondences
    it = GE.calculate_num_ransac_iterations(0.98, 8, 0.4)
    sizes = [int(v) for v in np.linspace(600, 2500, 16)]
    pairs = [two_view_correspondences(n, 100 + k, 0.4)[:2] for k, n in enumerate(sizes)]
    dev_pairs = [(torch.from_numpy(a.astype(np.float64)).cuda(), torch.from_numpy(b.astype(np.float64)).cuda()) for a, b in pairs]
    t0 = time.perf_counter()
    samples = [GE._samples_host(n, it, GE.RANSAC_SEED) for n in sizes]
    sampler_first_ms = (time.perf_counter() - t0) * 1e3
    dev_samples = [s.cuda() for s in samples]

    def resident():
        for (a, b), s in zip(dev_pairs, dev_samples):
            GE.ransac_device(a, b, it, samples=s)
    for _ in range(3):
        resident()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    N.profile_enable(True, local)
    e0.record()
    for _ in range(reps):
        resident()
    e1.record()
    torch.cuda.synchronize()
    kst = N.profile_collect(local)
    N.profile_enable(False, local)
    res_ms = e0.elapsed_time(e1) / reps
    evals = float(sum(sizes)) * it

    def host_call(threads):
        GE._samples_host.cache_clear()                     # every pair of a real run has its own size
        t = time.perf_counter()
        out = GE.find_inliers_many(pairs, max_iterations=it, threads=threads)
        return (time.perf_counter() - t) * 1e3, out
    host_call(8)
    ms1, _ = host_call(1)
    ms8, out8 = host_call(8)
    GE._samples_host.cache_clear()
    t0 = time.perf_counter()
    for n in sizes:
        GE._samples_host(n, it, GE.RANSAC_SEED)
    sampler_ms = (time.perf_counter() - t0) * 1e3

    # association: 2500 previous-frame matches against 2500 triangulated points; 2500 new 3-D points
    # against a store of 100 000
    rng = np.random.default_rng(0)
    prev = torch.from_numpy(rng.integers(0, 1920, (2500, 2)).astype(np.float64)).cuda()
    qry = torch.from_numpy(rng.integers(0, 1920, (2500, 2)).astype(np.float64)).cuda()
    store = torch.from_numpy(rng.normal(size=(100000, 3))).cuda()
    batch = torch.cat([store[torch.from_numpy(rng.integers(0, 100000, 1250)).cuda()], torch.from_numpy(rng.normal(size=(1250, 3))).cuda()])

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    assoc_ms = timed(lambda: AS.associate_device(prev, qry, 5.0))
    dedup_ms = timed(lambda: AS.dedup_device(batch, store))

    cpu_leg = None
    if cpu:
        from oracle import geometry as OG
        a, b = pairs[len(pairs) // 2]
        it_cpu = 300
        t0 = time.perf_counter()
        OG.find_inliers(a, b, max_iterations=it_cpu)
        dt = time.perf_counter() - t0
        cpu_leg = {"value": it_cpu * len(a) / dt, "unit": "hypothesis-correspondence evaluations/s", "cores": 1, "kind": "port",
                   "seconds": dt, "sample": f"{it_cpu} of the {it} hypotheses of one pair of {len(a)} correspondences through the "
                                            f"oracle (numpy/LAPACK restatement of SFM.py:126-160, bit-identical to the reference here)"}
    return {"workload": f"CameraPose.find_inliers, {it} hypotheses per pair, 16 pairs of 600..2500 correspondences (40% outliers)",
            "resident": {"ms_per_pair": res_ms / len(sizes), "pairs_per_s": len(sizes) / (res_ms * 1e-3),
                         "hypothesis_correspondence_evals_per_s": evals / (res_ms * 1e-3),
                         "kernels_ms_per_pair": {k: v[1] / reps / len(sizes) for k, v in sorted(kst.items())}},
            "host_call": {"ms_per_pair_1_thread": ms1 / len(sizes), "ms_per_pair_8_threads": ms8 / len(sizes),
                          "pairs_per_s_8_threads": len(sizes) / (ms8 * 1e-3),
                          "sampler_ms_per_pair_1_thread": sampler_ms / len(sizes), "sampler_first_call_ms": sampler_first_ms,
                          "inliers": [int(len(o[0])) for o in out8],
                          "note": "numpy arrays in, inlier arrays out; the time is the host replay of np.random.seed(5) / "
                                  "np.random.choice(n, 8, replace=False) (sequential MT19937 + rejection, ~1.5 ns per stream word)"},
            "association": {"associate_2500x2500_ms": assoc_ms, "dedup_2500_vs_100000_ms": dedup_ms},
            "cpu_baseline": cpu_leg}


# ---------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from sfmfromscratch_b200 import _native as N
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.extractor import extract_batch_device, make_params
    from sfmfromscratch_b200.matcher import match_batch_device
    from sfmfromscratch_b200.synth import synth_descriptor_base, synth_descriptors, synth_image

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.load_library()
    N.get_ctx(local)
    pk = peaks()

    # ---- inputs (synthetic, seeded; different per rank)
    base_imgs = np.stack([synth_image(IMG_H, IMG_W, 1000 * rank + s) for s in range(N_DISTINCT)])
    host_batch = torch.from_numpy(np.concatenate([base_imgs] * (BATCH // N_DISTINCT))).pin_memory()
    images = host_batch.to(dev)
    params, keep = make_params({}, pyramid=True)
    pairs_global = PL.consecutive_pairs(world * BATCH)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipe = PL.FeaturePipeline({}, RATIO, rank=rank, world=world)
    pipe.pair_block = 4
    # consecutive pairs (Runner.py:183-191): a pair is matched by the rank that owns its first image, and the
    # exchange is an all-gather of the ONE descriptor block per rank the neighbouring shard needs (PairPlan)
    plan = pipe.pair_plan(pairs_global, BATCH)
    my_pairs_np = plan.mine
    n_my_pairs = int(len(my_pairs_np))

    def step(imgs):
        # no host wait inside a step: the candidate-overflow flags accumulate on the device and are read
        # once after the timed region (pipe.overflow_since_last_check)
        out = pipe.extract(imgs, deferred_check=True)
        m = pipe.match_plan(plan, out['desc'], out['count'], cap=2500)
        return out, m

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- main timed region (resident inputs): K steps, clocks sampled live
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step(images)
    barrier()
    l0 = N.launch_count(local)
    sampler.start()
    total_ms = timed(lambda: step(images), args.steps, 0)
    if pipe.overflow_since_last_check():
        raise SystemExit("candidate buffer overflow in the timed region")
    sampler.stop_flag = True
    launches = N.launch_count(local) - l0
    sampler.join(timeout=2)
    ms_per_step = total_ms / args.steps
    pixels_per_step = world * BATCH * IMG_H * IMG_W
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- the same K steps again with the library's per-kernel CUDA events switched on (two event
    #      records per launch cost a few per cent, so `value` comes from the clean region above)
    N.profile_enable(True, local)
    prof_ms = timed(lambda: step(images), args.steps, 0) / args.steps
    kstat = N.profile_collect(local)
    N.profile_enable(False, local)

    # ---- roofline of the dominant kernel (k_harris)
    lp = level_pixels(IMG_H, IMG_W)
    kh = kstat.get("k_harris", (0, 0.0))
    kh_ms_step = kh[1] / args.steps if kh[1] else float("nan")
    # 4 B read + 4 B written per pyramid pixel, + 4 B per pixel of the next (exactly halved) level it emits
    harris_bytes = (8.0 * lp + 4.0 * (lp - IMG_H * IMG_W)) * BATCH
    achieved = harris_bytes / (kh_ms_step * 1e-3) / 1e9 if kh[1] else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("k_harris_dram_bytes_per_step")
    roofline = {"kernel": "k_harris (fused Sobel + second moments + 7x7 window + R)", "bound": "hbm",
                "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": (achieved / pk["hbm"]) if achieved else None,
                "traffic": traffic, "peak_source": pk["src"],
                "algorithmic_bytes_per_step": harris_bytes, "kernel_ms_per_step": kh_ms_step,
                "share_of_step": (kh_ms_step / prof_ms) if kh[1] else None,
                "note": "per-kernel CUDA events on the launching stream over a repeat of the timed region "
                        "(ms_per_step_profiled); the 147-FMA bit-exact window chain makes this kernel FP32-issue "
                        "bound (floor 0.35 ms/step at 128 FMA/clk/SM), see DESIGN.md section 6"}
    # the resource that actually bounds it: 147 order-preserving FMAs per pyramid pixel on the FP32 pipe
    fma_peak = 128.0 * 148 * 1.965e9 / 1e12                      # TFMA/s: 128 lanes/clk/SM x 148 SMs x max SM clock
    fma_ach = 147.0 * lp * BATCH / (kh_ms_step * 1e-3) / 1e12 if kh[1] else None
    roofline["fp32_pipe"] = {"fma_per_pyramid_pixel": 147, "achieved_tfma_per_s": fma_ach, "peak_tfma_per_s": fma_peak,
                             "frac": (fma_ach / fma_peak) if fma_ach else None,
                             "window_stage_alone_frac": 0.72,
                             "note": "window chains only (Sobel/products add 30 FP32 ops per pixel on the same pipe); "
                                     "window_stage_alone_frac = scripts/micro/window_rate.cu, the stage's ceiling at any occupancy"}
    kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in sorted(kstat.items())}

    # ---- e2e: host buffers in and out, copies inside the timed region
    cap = 2500
    h_x = torch.empty((BATCH, cap), dtype=torch.int32).pin_memory()
    h_y = torch.empty((BATCH, cap), dtype=torch.int32).pin_memory()
    h_d = torch.empty((BATCH, cap, 128), dtype=torch.float32).pin_memory()
    h_c = torch.empty((BATCH,), dtype=torch.int32).pin_memory()
    h_m = torch.empty((max(n_my_pairs, 1), cap, 2), dtype=torch.int32).pin_memory()
    h_mc = torch.empty((max(n_my_pairs, 1), cap), dtype=torch.float32).pin_memory()
    h_mn = torch.empty((max(n_my_pairs, 1),), dtype=torch.int32).pin_memory()

    host_out = {'x': h_x, 'y': h_y, 'desc': h_d, 'count': h_c, 'matches': h_m, 'conf': h_mc, 'mcount': h_mn}
    host_out2 = {k: torch.empty_like(v).pin_memory() for k, v in host_out.items()}     # consecutive steps land in alternate sets
    flip = [0]

    def e2e_step():
        flip[0] ^= 1
        pipe.stream_host(host_batch, pairs_global, host_out2 if flip[0] else host_out, chunk=8)

    def e2e_single():
        pipe.run_host(host_batch, pairs_global, host_out, chunk=8)

    e2e_steps = max(3, min(args.steps, 30))
    e2e_ms = timed(e2e_step, e2e_steps, min(args.warmup, 3)) / e2e_steps
    if not pipe.drain():
        raise SystemExit("candidate buffer overflow in the e2e leg")
    e2e_single_ms = timed(e2e_single, max(3, e2e_steps // 3), 1) / max(3, e2e_steps // 3)
    # the PCIe floor of that step: the same pinned host batch copied to the device with nothing else running
    scratch = torch.empty_like(images)
    h2d_ms = timed(lambda: scratch.copy_(host_batch, non_blocking=True), 10, 2) / 10
    del scratch
    h2d = host_batch.numel() * 4
    d2h = (h_x.numel() + h_y.numel() + h_d.numel() + h_c.numel()) * 4
    if n_my_pairs:
        d2h += (h_m.numel() + h_mc.numel() + h_mn.numel()) * 4
    e2e = {"value": pixels_per_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "ms_per_step_one_batch_at_a_time": e2e_single_ms,
           "h2d_copy_alone_ms": h2d_ms, "h2d_copy_alone_gb_per_s": host_batch.numel() * 4 / (h2d_ms * 1e-3) / 1e9,
           "api": "FeaturePipeline.stream_host: pinned host images in, pinned host keypoints/descriptors/matches out, every "
                  "step; chunks of 8 images, H2D / kernels / D2H on three streams, no host wait between steps so the copy of "
                  "step k+1 runs under the matching and read-back of step k (run_host, which waits for each batch, is "
                  "ms_per_step_one_batch_at_a_time)"}

    # ---- matcher alone on configs[4]-shaped pairs
    base = synth_descriptor_base(MATCH_N)
    sets = np.stack([synth_descriptors(MATCH_N, 100 * rank + i, base=base) for i in range(MATCH_SETS)])
    d_sets = torch.from_numpy(sets).to(dev)
    d_cnt = torch.full((MATCH_SETS,), MATCH_N, dtype=torch.int32, device=dev)
    d_pairs = torch.from_numpy(PL.all_pairs(MATCH_SETS)).to(dev)
    n_mp = int(d_pairs.shape[0])
    msteps = max(3, min(args.steps, 20))
    for _ in range(3):
        match_batch_device(d_sets, d_cnt, d_pairs, RATIO, cap=MATCH_N)
    barrier()
    N.profile_enable(True, local)
    m_ms = timed(lambda: match_batch_device(d_sets, d_cnt, d_pairs, RATIO, cap=MATCH_N), msteps, 0) / msteps
    mstat = N.profile_collect(local)
    N.profile_enable(False, local)
    mm, mc, mcnt, mst = match_batch_device(d_sets, d_cnt, d_pairs, RATIO, cap=MATCH_N, want_stats=True)
    torch.cuda.synchronize()
    dpairs = float(n_mp) * MATCH_N * MATCH_N
    tc = mstat.get("k_match_tc", (0, 0.0))
    tc_ms = tc[1] / msteps if tc[1] else float("nan")
    tc_ach = 256.0 * dpairs / (tc_ms * 1e-3) / 1e12 if tc[1] else None
    match = {"metric": "nn_ratio_descriptor_pairs_per_s", "value": world * dpairs / (m_ms * 1e-3), "unit": "descriptor-pairs/s",
             "image_pairs_per_s": world * n_mp / (m_ms * 1e-3), "ms_per_step": m_ms, "steps": msteps,
             "workload": f"all {n_mp} pairs of {MATCH_SETS} sets of {MATCH_N} x 128 f32 descriptors per GPU (configs[4] pair shape), thr {RATIO}",
             "matches_per_pair_mean": float(mcnt.float().mean().item()),
             "rows_rescanned_frac": float(mst[:, 0].float().sum().item()) / (n_mp * MATCH_N),
             "candidate_groups_rechecked_per_row": float(mst[:, 1].float().sum().item()) / (n_mp * MATCH_N),
             "whole_path": {"achieved": 256.0 * dpairs / (m_ms * 1e-3) / 1e12, "unit": "TFLOP/s (algorithmic, every matcher kernel)",
                            "frac_of_sustained_peak": 256.0 * dpairs / (m_ms * 1e-3) / 1e12 / pk["tf_sust"]},
             "kernels": {k: {"launches_per_step": v[0] / msteps, "ms_per_step": v[1] / msteps} for k, v in sorted(mstat.items())},
             "roofline": {"kernel": "k_match_tc (tcgen05 M128xN256xK16 fp16 GEMM, fused key + top-4 groups-of-8 epilogue)", "bound": "tensor",
                          "achieved": tc_ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                          "frac": (tc_ach / pk["tf_sust"]) if tc_ach else None, "traffic": None,
                          "peak_burst": pk["tf_burst"], "frac_of_burst": (tc_ach / pk["tf_burst"]) if tc_ach else None,
                          "peak_source": pk["src"] + ", sustained bf16 (fp16 runs at the same rate)",
                          "flops": "256 per descriptor pair (algorithmic == executed: single fp16 pass)"}}

    # ---- configs[1] literally: ONE 1080p image (latency-bound: 47 MB of traffic, ~25 launches)
    single = None
    if rank == 0:
        from sfmfromscratch_b200 import ScaleRotInvSIFT
        one = images[:1]
        for _ in range(3):
            extract_batch_device(one, params, want_aux=False)
        torch.cuda.synchronize()
        a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a1.record()
        for _ in range(20):
            extract_batch_device(one, params, want_aux=False)
        b1.record()
        torch.cuda.synchronize()
        lat = a1.elapsed_time(b1) / 20
        ge = PL.GraphedExtractor(one, {})
        for _ in range(3):
            ge.run(one)
        torch.cuda.synchronize()
        a1.record()
        for _ in range(20):
            ge.run(one)
        b1.record()
        torch.cuda.synchronize()
        lat_graph = a1.elapsed_time(b1) / 20
        gb = PL.GraphedExtractor(images, {})
        for _ in range(3):
            gb.run(images)
        torch.cuda.synchronize()
        a1.record()
        for _ in range(20):
            gb.run(images)
        b1.record()
        torch.cuda.synchronize()
        batch_graph = a1.elapsed_time(b1) / 20
        del ge, gb
        img_np = base_imgs[0]
        ScaleRotInvSIFT(img_np, {})
        t0 = time.time()
        for _ in range(5):
            e = ScaleRotInvSIFT(img_np, {})
        host_ms = (time.time() - t0) / 5 * 1e3
        single = {"workload": "one 1920x1080 image, ScaleRotInvSIFT defaults (configs[1])",
                  "resident_ms": lat, "resident_mpixel_per_s": IMG_H * IMG_W / (lat * 1e-3) / 1e6,
                  "resident_cuda_graph_ms": lat_graph,
                  "batch32_extract_only_cuda_graph_ms": batch_graph,
                  "class_call_ms": host_ms, "class_call_mpixel_per_s": IMG_H * IMG_W / (host_ms * 1e-3) / 1e6,
                  "keypoints": int(len(e.detect_keypoints()[0])),
                  "note": "class_call = ScaleRotInvSIFT(numpy image, {}) -> numpy keypoints/descriptors, pageable host memory"}

    # ---- configs[2]: one 3840x2160 image at ~20 k keypoints, matched against its second view
    cfg2 = None
    if rank == 0 and not args.no_4k:
        from sfmfromscratch_b200.synth import second_view
        from sfmfromscratch_b200.matcher import match_device
        a4 = synth_image(2160, 3840, 5)
        b4 = second_view(a4, 6)
        d4 = torch.from_numpy(np.stack([a4, b4])).to(dev)
        p4, keep4 = make_params({'num_interest_points': 32000}, pyramid=True)

        def run4():
            o = extract_batch_device(d4, p4, want_aux=False)
            n0, n1 = (int(v) for v in o['count'].cpu())
            m = match_device(o['desc'][0, :n0], o['desc'][1, :n1], RATIO)
            return n0, n1, m

        for _ in range(2):
            run4()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            n0, n1, m4 = run4()
        torch.cuda.synchronize()
        ms4 = (time.time() - t0) / 5 * 1e3
        cfg2 = {"workload": "two 3840x2160 views, num_interest_points 32000: extraction of both + NN-ratio matching (configs[2])",
                "ms": ms4, "keypoints": [n0, n1], "matches": int(m4[2].cpu()[0]),
                "mpixel_per_s": 2 * 2160 * 3840 / (ms4 * 1e-3) / 1e6,
                "note": "wall clock incl. the host read of the keypoint counts between extraction and matching"}
        del keep4

    # ---- two-view RANSAC and point association on the matcher's output
    geom = None
    if rank == 0 and not args.no_geometry:
        geom = geometry_leg(torch, N, local, cpu=(world == 1 and not args.no_cpu))

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        n_cpu = 8
        v, dt = cpu_sample(cpu_images(n_cpu))
        cpu = {"value": v, "unit": "Mpixel/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"{n_cpu} of the step's {BATCH} 1080p images + their {n_cpu - 1} consecutive-pair matches through the "
                         f"CPU oracle (C + numpy port of the reference), one process; the reference's own Python loops are ~15x "
                         f"slower per image (BASELINE.md section 2)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "images_per_gpu": BATCH, "image": [IMG_H, IMG_W],
                           "distinct_images": N_DISTINCT, "pairs_per_gpu": n_my_pairs,
                           "l2": "inputs larger than L2 (265 MB of images, 352 MB of R planes per step)",
                           "parallelism": f"image shards x{world}, all-gather (NCCL) of the {plan.K} descriptor block(s) per rank that "
                                          f"another shard's pairs need (policy {plan.policy}), pair shards"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": sampler.summary(), "ms_per_step_profiled": prof_ms, "kernels": kernels, "match": match,
                "single_image": single, "config2_4k_pair": cfg2, "geometry": geom}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    del keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-4k", action="store_true", help="skip the configs[2] leg")
    ap.add_argument("--no-geometry", action="store_true", help="skip the RANSAC / association leg (SURVEY 8f rows 2-3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
