#!/usr/bin/env python
"""bench.py -- throughput of the SfmFromScratch feature hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is BASELINE.json configs[3] as written: SIFT extraction (ScaleRotInvSIFT, reference
defaults) of a 256-frame synthetic 1920x1080 sequence, sharded image-wise over the N GPUs (256 / N
frames per GPU, extracted in batches of 32), the exchange of the descriptor blocks the pair list
needs when N > 1, and NN-ratio matching of this rank's share of the 255 consecutive frame pairs
(what Runner.py:183-191 matches).  Total work is fixed ("scaling": "strong").  `value` is input
Mpixel/s of the whole job with the frames resident in HBM; `e2e` is the same job through host
buffers (pinned H2D of the frames, D2H of keypoints, descriptors and matches inside the timed
region).  Further legs on the same line:

  sustained      the step repeated for >= 2 s, SM clock under load
  all_pairs      configs[4] as written: 512 images x 8192 descriptors, image shards, the full NCCL
                 all-gather INSIDE the timed region, 130 816 pairs dealt block-cyclically
  match          matcher alone on 66 pairs of 8192 x 8192 (per-kernel times, tensor-pipe roofline)
  single_image   configs[1]: one 1080p image
  config0_two_view   configs[0]: 640x480 two-view chain (extract x2 + match + coords + find_inliers)
  config2_4k_pair    configs[2]: two 3840x2160 views at ~20 k keypoints + matching
  geometry       SURVEY 8f rows 2-3

The reference arm (--impl reference) times the reference's own classes (staged unmodified into
oracle/_ref by oracle/stage_reference.py) on all host cores, with the oracle port beside it.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W = 1080, 1920
N_IMAGES = 256                  # configs[3]
BATCH = 32                      # frames per extraction call
AP_IMAGES, AP_N, AP_CHUNK = 512, 8192, 256      # configs[4]; pairs per matcher call
MATCH_SETS, MATCH_N = 12, 8192  # matcher-only leg: all 66 pairs of 12 sets of 8192 descriptors
RATIO = 0.8
METRIC = "sift_extract_plus_nn_ratio_match_input_mpixel_per_s"
WORKLOAD = (f"configs[3]: ScaleRotInvSIFT(defaults: 4 levels, k=2500) on a {N_IMAGES}-frame synthetic {IMG_W}x{IMG_H} f32 "
            f"sequence sharded image-wise over the GPUs + NN-ratio matching (thr {RATIO}) of the consecutive frame pairs")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def ncu_traffic():
    """dram bytes per launch of the dominant kernels, from the round's `ncu --set full` capture (profiles/)."""
    for name in ("r02_ncu_traffic.json", "ncu_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            d["_file"] = "profiles/" + name
            return d
    return {}


def level_pixels(h, w, levels=4, f=2):
    tot = 0
    for _ in range(levels):
        tot += h * w
        h, w = int(h / f), int(w / f)
    return tot


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during a timed region."""

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

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_min_mhz": (s[0] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------- CPU arm

def cpu_frames(n_images):
    """The first n frames of the bench sequence (the same bytes the GPU arm's rank 0 extracts)."""
    from sfmfromscratch_b200.synth import frame_sequence
    return list(frame_sequence(IMG_H, IMG_W, 0, n_images, N_IMAGES, threads=min(8, os.cpu_count() or 1)))


def _port_init():
    from oracle import oracle as O
    O.lib()


def _port_extract(img):
    from oracle import oracle as O
    return O.ScaleRotInvSIFT(img, {}).extract_descriptors()


def _port_match(args):
    from oracle import oracle as O
    return len(O.NNRatioFeatureMatcher(RATIO).match_features_ratio_test(*args)[0])


def _ref_init(ref_dir):
    sys.path.insert(0, ref_dir)
    try:
        import cv2
        cv2.setNumThreads(1)          # one worker process per core: no nested threading
    except Exception:
        pass
    import FeatureExtractor, FeatureMatcher      # noqa: F401  (the staged, unmodified reference)
    assert os.path.abspath(FeatureExtractor.__file__).startswith(os.path.abspath(ref_dir))


def _ref_extract(img):
    from FeatureExtractor.SIFT.ScaleRotInvSIFT import ScaleRotInvSIFT     # ScaleRotInvSIFT.py:9-16: all work in __init__
    e = ScaleRotInvSIFT(img, {})
    e.detect_keypoints()
    return e.extract_descriptors()


def _ref_match(args):
    from FeatureMatcher import NNRatioFeatureMatcher
    return len(NNRatioFeatureMatcher(ratio_threshold=RATIO).match_features_ratio_test(*args)[0])


def cpu_sample(images, extract, match, pool=None, match_pool=None):
    """`extract` on every image + `match` on the consecutive pairs, in this process or over `pool`; returns
    (Mpixel/s, seconds).  Input synthesis and worker start-up are outside the timed region."""
    n = len(images)
    t0 = time.time()
    if pool is None:
        feats = [extract(im) for im in images]
        for i in range(n - 1):
            match((feats[i], feats[i + 1]))
    else:
        feats = pool.map(extract, images, chunksize=1)
        (match_pool or pool).map(match, [(feats[i], feats[i + 1]) for i in range(n - 1)], chunksize=1)
    dt = time.time() - t0
    return n * IMG_H * IMG_W / dt / 1e6, dt


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: every core extracts one
    frame of the workload's sequence through the UNMODIFIED ScaleRotInvSIFT, then the consecutive pairs go
    through the unmodified NNRatioFeatureMatcher (kind "reference"); the oracle port is timed beside it.  When
    no staged reference is present (oracle/_ref), the port is the line's value (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    from oracle import stage_reference
    O.build()
    ref_dir = stage_reference.stage() or stage_reference.staged()
    if args.port_only:
        ref_dir = None
    cores = os.cpu_count() or 1
    n = max(cores, 2)
    images = cpu_frames(min(n, N_IMAGES))
    n = len(images)
    steps = max(1, min(args.steps, 2 if ref_dir else 3))       # bounded: each step is a sample of the workload
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores, initializer=_port_init) as pool:
        pool.map(_port_extract, images[:cores], chunksize=1)    # warm-up: imports, page-in
        pv, pt = zip(*[cpu_sample(images, _port_extract, _port_match, pool) for _ in range(3 if ref_dir else steps)])
    port = {"value": float(np.mean(pv)), "unit": "Mpixel/s", "cores": cores, "kind": "port", "ms_per_step": float(np.mean(pt)) * 1e3,
            "sample": f"{n} frames + {n - 1} consecutive pairs per step through the oracle port (C + numpy restatement), "
                      f"{cores} worker processes"}
    if ref_dir:
        # the reference matcher materialises (n1, n2, 128) float32 temporaries twice (~3 GB per 1.7 k x 1.7 k pair):
        # at most 8 pairs in flight
        with ctx.Pool(cores, initializer=_ref_init, initargs=(ref_dir,)) as pool, \
                ctx.Pool(min(cores, 8), initializer=_ref_init, initargs=(ref_dir,)) as mpool:
            pool.map(_ref_extract, [im[:270, :480].copy() for im in images[:cores]], chunksize=1)   # warm-up: imports
            vals, times = zip(*[cpu_sample(images, _ref_extract, _ref_match, pool, mpool) for _ in range(steps)])
        kind, value, ms = "reference", float(np.mean(vals)), float(np.mean(times)) * 1e3
        sample = (f"each step = {n} of the workload's {N_IMAGES} 1080p frames (one per core) + their {n - 1} consecutive-pair matches "
                  f"through the UNMODIFIED reference classes (FeatureExtractor/SIFT/ScaleRotInvSIFT.py, "
                  f"FeatureMatcher/NNRatioFeatureMatcher.py, staged into oracle/_ref), {cores} worker processes "
                  f"(matching: {min(cores, 8)}, ~3 GB of temporaries per pair), cv2.setNumThreads(1); {steps} steps timed")
    else:
        kind, value, ms, sample = "port", port["value"], port["ms_per_step"], port["sample"] + f"; {steps} steps timed"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": kind, "sample": sample},
            "port": port,
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

    def host_call(threads, cold):
        if cold:
            GE._samples_host.cache_clear()                 # every pair of a first run has its own size
        t = time.perf_counter()
        out = GE.find_inliers_many(pairs, max_iterations=it, threads=threads)
        return (time.perf_counter() - t) * 1e3, out
    host_call(8, True)
    ms1, _ = host_call(1, True)
    ms8, out8 = host_call(8, True)
    ms8w, _ = host_call(8, False)                          # 8-subset tables of these sizes already drawn (per-process cache)
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
                          "ms_per_pair_8_threads_tables_cached": ms8w / len(sizes),
                          "pairs_per_s_8_threads": len(sizes) / (ms8 * 1e-3),
                          "sampler_ms_per_pair_1_thread": sampler_ms / len(sizes), "sampler_first_call_ms": sampler_first_ms,
                          "inliers": [int(len(o[0])) for o in out8],
                          "note": "numpy arrays in, inlier arrays out; the cold time is the host replay of np.random.seed(5) / "
                                  "np.random.choice(n, 8, replace=False) (sequential MT19937 + rejection, ~1.5 ns per stream word); "
                                  "the draws depend on (n, iterations) only, so a table is drawn once per size and process"},
            "association": {"associate_2500x2500_ms": assoc_ms, "dedup_2500_vs_100000_ms": dedup_ms},
            "cpu_baseline": cpu_leg}


# ---------------------------------------------------------------------------- configs[4]

def synth_descriptor_block(torch, n_img, n, seed, dev):
    """RootSIFT-shaped descriptors generated on the device (numpy synthesis of 512 x 8192 rows takes minutes):
    half of each image's rows are noisy copies of rows of a shared base set, so true matches exist between any
    two images."""
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
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


def all_pairs_leg(torch, dist, PL, rank, world, local, dev, pk, n_images=AP_IMAGES, n=AP_N, chunk=AP_CHUNK, reps=1):
    """BASELINE.json configs[4] as written, strong scaling: every rank holds the descriptor blocks of its
    contiguous shard of the 512 images; inside the timed region: ONE all-gather of the blocks (NCCL / NVLink),
    the per-set preparation (fp16 copy + norms, once), and this rank's block-cyclic share of the 130 816 pairs
    in chunks of 256 (SFM_MATCH_PREPARED).  CUDA events, max over ranks."""
    from sfmfromscratch_b200.matcher import match_batch_device, match_workspace
    per = (n_images + world - 1) // world
    s0, s1 = PL.shard_images(n_images, rank, world)
    mine = torch.zeros((per, n, 128), device=dev)
    mine[: s1 - s0] = synth_descriptor_block(torch, s1 - s0, n, 77 + rank, dev)
    counts = torch.zeros((per,), dtype=torch.int32, device=dev)
    counts[: s1 - s0] = n
    pairs = PL.deal_pairs(PL.all_pairs(n_images), rank, world)
    pairs_dev = torch.from_numpy(np.ascontiguousarray(pairs)).to(dev)
    ws = match_workspace(per * world, n, chunk, dev)

    def job():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        total = torch.zeros((), dtype=torch.int64, device=dev)
        e0.record()
        desc_all, counts_all = PL.gather_descriptors(mine, counts)
        e1.record()
        for c0 in range(0, len(pairs), chunk):
            m, c, cnt = match_batch_device(desc_all, counts_all, pairs_dev[c0:c0 + chunk], RATIO, cap=n, ws=ws, prepared=c0 > 0)
            total += cnt.sum()          # stays on the device: no host sync per chunk
        e2.record()
        return e0, e1, e2, total

    # warm-up: one chunk of local pairs (kernel modules, tensor-map entry point, allocator blocks)
    wp = torch.from_numpy(np.ascontiguousarray(PL.all_pairs(min(per, 24))[:chunk])).to(dev)
    if len(wp):
        match_batch_device(mine, counts.clamp(min=2), wp, RATIO, cap=n, ws=match_workspace(per, n, chunk, dev))
    torch.cuda.synchronize()
    best = None
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2, total = job()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], device=dev)
        tm = total.to(torch.float64).reshape(1)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tm)
        if best is None or ms[0].item() < best[0]:
            best = (ms[0].item(), ms[1].item(), int(tm.item()))
    clocks = sampler.finish()
    total_ms, ag_ms, matches = best
    npairs = n_images * (n_images - 1) // 2
    dpairs = float(npairs) * n * n
    tf = 256.0 * dpairs / (total_ms * 1e-3) / 1e12
    return {"workload": f"configs[4]: all-pairs NN-ratio matching over {n_images} images x {n} x 128 f32 descriptors "
                        f"({npairs} image pairs), image shards x{world}, one all-gather, pairs dealt block-cyclically in blocks of 64",
            "metric": "nn_ratio_descriptor_pairs_per_s", "value": dpairs / (total_ms * 1e-3), "unit": "descriptor-pairs/s",
            "image_pairs_per_s": npairs / (total_ms * 1e-3), "seconds": total_ms * 1e-3, "all_gather_ms": ag_ms,
            "all_gather_bytes_per_rank": int(per * n * 128 * 4) if world > 1 else 0,
            "scaling": "strong", "n_gpus": world, "reps": reps, "pairs_per_rank": int(len(pairs)), "chunk_pairs": chunk,
            "matches_total": matches,
            "whole_path": {"achieved": tf, "unit": "TFLOP/s (algorithmic, 256 flop per descriptor pair, every kernel and the all-gather, all GPUs)",
                           "per_gpu": tf / world, "frac_of_sustained_peak": tf / world / pk["tf_sust"],
                           "frac_of_burst_peak": tf / world / pk["tf_burst"]},
            "clocks": clocks, "timing": "CUDA events around all-gather + every chunk, max over ranks, after one warm-up chunk"}


# ---------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from sfmfromscratch_b200 import _native as N
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.extractor import extract_batch_device, make_params
    from sfmfromscratch_b200.matcher import match_batch_device
    from sfmfromscratch_b200.synth import frame_sequence, synth_descriptor_base, synth_descriptors, synth_image

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
    traffic = ncu_traffic()

    # ---- inputs: this rank's contiguous shard of the 256-frame sequence (synthetic, seeded)
    per = (N_IMAGES + world - 1) // world
    n_job = per * world                                  # == 256 for 1, 2, 4, 8 GPUs
    batch = min(BATCH, per)
    host_shard = torch.from_numpy(frame_sequence(IMG_H, IMG_W, rank * per, per, n_job,
                                                 threads=max(1, min(16, (os.cpu_count() or 1) // max(1, world))))).pin_memory()
    images = host_shard.to(dev)
    params, keep = make_params({}, pyramid=True)
    pairs_global = PL.consecutive_pairs(n_job)
    cap = 2500

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipe = PL.FeaturePipeline({}, RATIO, rank=rank, world=world)
    # consecutive pairs (Runner.py:183-191): a pair is matched by the rank that owns its first image, and the
    # exchange is an all-gather of the ONE descriptor block per rank the neighbouring shard needs (PairPlan)
    plan = pipe.pair_plan(pairs_global, per)
    n_my_pairs = int(len(plan.mine))
    # this rank's result tables; the descriptor table carries the exchange's slots behind the rank's own blocks
    shard = pipe.shard_tables(plan, cap, dev)
    bounds = [(b0, min(b0 + batch, per)) for b0 in range(0, per, batch)]

    last = {}

    def step():
        # FeaturePipeline.stream_resident: extraction in calls of `batch` frames on the current stream, exchange +
        # matching on a second stream (under the next step's extraction), two sets of result tables used alternately.
        # No host wait inside a step: the candidate-overflow flags accumulate on the device and are read once after
        # the timed region (pipe.overflow_since_last_check).  The timed region ends with a device-wide synchronize, so
        # every step's matching is inside it.
        if args.serial_step:
            return step_serial()
        tabs, m, ev = pipe.stream_resident(images, plan, batch=batch, cap=cap)
        last['tabs'], last['ev'] = tabs, ev
        return m

    def step_serial():
        # the same work on ONE stream (matching behind the extraction): what the per-kernel pass times, because a
        # kernel's event interval means nothing while another stream's kernels share the device
        for b0, b1 in bounds:
            pipe.extract(images[b0:b1], deferred_check=True, out={k: v[b0:b1] for k, v in shard.items()})
        last['tabs'] = shard
        return pipe.match_plan(plan, shard['desc'], shard['count'], cap=cap)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        pipe.join_resident()          # the closing event is recorded behind the matcher stream of stream_resident as well
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- main timed region (resident inputs): K steps, clocks sampled live
    for _ in range(args.warmup):
        step()
    barrier()
    l0 = N.launch_count(local)
    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed(step, args.steps, 0)
    clocks = sampler.finish()
    if pipe.overflow_since_last_check():
        raise SystemExit("candidate buffer overflow in the timed region")
    launches = N.launch_count(local) - l0
    ms_per_step = total_ms / args.steps
    pixels_per_step = n_job * IMG_H * IMG_W
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6
    mcount = step()
    torch.cuda.synchronize()
    matches_per_pair = float(mcount[2].float().mean().item()) if mcount is not None else None
    kp_per_image = float(last['tabs']['count'][:per].float().mean().item())

    # ---- the same step for >= 2 s: what the clocks do under seconds of load
    n_sust = max(args.steps, int(math.ceil(2200.0 / ms_per_step)))
    s2 = ClockSampler(local)
    s2.start()
    sust_ms = timed(step, n_sust, 0)
    sustained = {"seconds": sust_ms * 1e-3, "steps": n_sust, "ms_per_step": sust_ms / n_sust,
                 "value": pixels_per_step / (sust_ms / n_sust * 1e-3) / 1e6, "unit": "Mpixel/s", "clocks": s2.finish()}

    # ---- the same K steps again with the library's per-kernel CUDA events switched on (two event
    #      records per launch cost a few per cent, so `value` comes from the clean region above)
    #      -- on one stream (step_serial), so that every kernel's event interval is its own
    prof_steps = min(args.steps, 50)
    serial_ms = timed(step_serial, prof_steps, 2) / prof_steps
    N.profile_enable(True, local)
    prof_ms = timed(step_serial, prof_steps, 0) / prof_steps
    kstat = N.profile_collect(local)
    N.profile_enable(False, local)
    kernels = {k: {"launches_per_step": v[0] / prof_steps, "ms_per_step": v[1] / prof_steps} for k, v in sorted(kstat.items())}

    # ---- the exchange alone (N > 1): all-gather of the K blocks per rank + counts
    exchange_ms = None
    if world > 1:
        exchange_ms = timed(lambda: PL.exchange_for(plan, shard['desc'], shard['count']), 50, 5) / 50

    # ---- roofline of the dominant kernel (k_harris): FP32-pipe bound by construction (147 order-preserving
    #      FMAs per pyramid pixel), HBM beside it
    lp = level_pixels(IMG_H, IMG_W)
    kh = tuple(a + b for a, b in zip(kstat.get("k_harris", (0, 0.0)), kstat.get("k_harris_stream", (0, 0.0))))   # both Harris kernels
    kh_ms_step = kh[1] / prof_steps if kh[1] else float("nan")
    # 4 B read + 4 B written per pyramid pixel, + 4 B per pixel of the next (exactly halved) level it emits
    harris_bytes = (8.0 * lp + 4.0 * (lp - IMG_H * IMG_W)) * per
    hbm_ach = harris_bytes / (kh_ms_step * 1e-3) / 1e9 if kh[1] else None
    fp32_peak = 2.0 * 128.0 * 148 * 1.965e9 / 1e12               # TFLOP/s: 128 FMA lanes/clk/SM x 148 SMs x max SM clock
    fp32_ach = 2.0 * 147.0 * lp * per / (kh_ms_step * 1e-3) / 1e12 if kh[1] else None
    th = traffic.get("k_harris_dram_bytes_per_image")
    roofline = {"kernel": "k_harris: k_harris_stream<7> (persistent warp-specialised: TMA tiles -> producer warps -> shared-memory ring of "
                          "product rows -> consumer warps, 4 output rows per thread) on levels 0-2 + the tile kernel on level 3; "
                          "fused Sobel + second moments + 7x7 window + R + histogram + next pyramid level; every launch of a step",
                "bound": "fp32", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": (fp32_ach / fp32_peak) if fp32_ach else None,
                "traffic": (th * per) if th else None, "traffic_source": traffic.get("_file"),
                "peak_source": "2 x 128 FMA/clk/SM x 148 SMs x 1.965 GHz (FP32 pipe; no measured FP32 figure in MEASURED_PEAKS.json)",
                "algorithmic_flops_per_step": 2.0 * 147.0 * lp * per,
                "hbm": {"achieved": hbm_ach, "peak": pk["hbm"], "unit": "GB/s", "frac": (hbm_ach / pk["hbm"]) if hbm_ach else None,
                        "algorithmic_bytes_per_step": harris_bytes, "peak_source": pk["src"]},
                "kernel_ms_per_step": kh_ms_step, "share_of_step": (kh_ms_step / prof_ms) if kh[1] else None,
                "window_stage_alone_frac": 0.85,
                "note": "per-kernel CUDA events on the launching stream over a repeat of the timed region; the bit-exact 3 x 49 "
                        "dependent fmaf chains per pixel make this kernel FP32-issue bound (floor = frac 1.0), which caps its HBM "
                        "use near 0.35; window_stage_alone_frac = scripts/micro/window_forms.cu, the 4-row window stage alone "
                        "at 2 warps per scheduler (0.72 for the 2-row stage of the tile kernel)"}
    # the streaming pass over R (k_nms: window maxima + median-bucket compaction): HBM-bound
    kn = kstat.get("k_nms", (0, 0.0))
    kn_ms = kn[1] / prof_steps if kn[1] else float("nan")
    nms_bytes = 4.0 * lp * per
    tn = traffic.get("k_nms_dram_bytes_per_image")
    roofline_nms = {"kernel": "k_nms (one pass over R: window maxima + median-bucket compaction; all pyramid levels in one launch per extraction call)",
                    "bound": "hbm", "achieved": nms_bytes / (kn_ms * 1e-3) / 1e9 if kn[1] else None, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": nms_bytes / (kn_ms * 1e-3) / 1e9 / pk["hbm"] if kn[1] else None,
                    "traffic": (tn * per) if tn else None, "algorithmic_bytes_per_step": nms_bytes, "kernel_ms_per_step": kn_ms}
    # whole extraction against SURVEY 8d's per-pixel budget (16 B per pyramid pixel + 4 B per next-level pixel)
    ext_ms = sum(v[1] for k, v in kstat.items() if not k.startswith("k_match") and k != "k_work_scan") / prof_steps
    ext_bytes = (16.0 * lp + 4.0 * (lp - IMG_H * IMG_W)) * per
    extraction = {"ms_per_step": ext_ms, "algorithmic_bytes_per_step": ext_bytes,
                  "hbm_frac": ext_bytes / (ext_ms * 1e-3) / 1e9 / pk["hbm"] if ext_ms else None,
                  "non_harris_ms_per_step": ext_ms - (kh_ms_step if kh[1] else 0.0)}

    # ---- e2e: host buffers in and out, copies inside the timed region
    n_mp = max(n_my_pairs, 1)
    def host_outs():
        return {'x': torch.empty((per, cap), dtype=torch.int32).pin_memory(), 'y': torch.empty((per, cap), dtype=torch.int32).pin_memory(),
                'desc': torch.empty((per, cap, 128), dtype=torch.float32).pin_memory(),
                'count': torch.empty((per,), dtype=torch.int32).pin_memory(),
                'matches': torch.empty((n_mp, cap, 2), dtype=torch.int32).pin_memory(),
                'conf': torch.empty((n_mp, cap), dtype=torch.float32).pin_memory(),
                'mcount': torch.empty((n_mp,), dtype=torch.int32).pin_memory()}
    host_out, host_out2 = host_outs(), host_outs()           # consecutive steps land in alternate sets
    flip = [0]

    def e2e_step():
        flip[0] ^= 1
        pipe.stream_host(host_shard, pairs_global, host_out2 if flip[0] else host_out, chunk=8)

    e2e_steps = max(3, min(args.steps, int(math.ceil(1500.0 / (3.0 * ms_per_step)))))
    e2e_ms = timed(e2e_step, e2e_steps, 2) / e2e_steps
    if not pipe.drain():
        raise SystemExit("candidate buffer overflow in the e2e leg")
    single_steps = max(2, e2e_steps // 3)
    e2e_single_ms = timed(lambda: pipe.run_host(host_shard, pairs_global, host_out, chunk=8), single_steps, 1) / single_steps
    # the PCIe floor of that step: the same pinned host shard copied to the device with nothing else running
    scratch = torch.empty_like(images)
    h2d_ms = timed(lambda: scratch.copy_(host_shard, non_blocking=True), 5, 2) / 5
    del scratch
    h2d = host_shard.numel() * 4
    d2h = sum(host_out[k].numel() for k in ('x', 'y', 'desc', 'count')) * 4
    if n_my_pairs:
        d2h += sum(host_out[k].numel() for k in ('matches', 'conf', 'mcount')) * 4
    e2e = {"value": pixels_per_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "ms_per_step_one_job_at_a_time": e2e_single_ms,
           "h2d_copy_alone_ms": h2d_ms, "h2d_copy_alone_gb_per_s": h2d / (h2d_ms * 1e-3) / 1e9,
           "host_link_ceiling_mpixel_per_s": pixels_per_step / (h2d_ms * 1e-3) / 1e6,
           "api": "FeaturePipeline.stream_host: this rank's pinned host frames in, pinned host keypoints/descriptors/matches out, "
                  "every step; chunks of 8 frames, H2D / kernels / D2H on three streams, no host wait between steps so the copy of "
                  "step k+1 runs under the matching and read-back of step k (run_host, which waits for each job, is "
                  "ms_per_step_one_job_at_a_time).  h2d/d2h bytes are per rank; host_link_ceiling = the job's pixels over the time "
                  "the bare H2D copy of the frames takes with all ranks copying at once"}
    del host_out, host_out2

    # ---- matcher alone on configs[4]-shaped pairs (per-kernel times)
    base = synth_descriptor_base(MATCH_N)
    sets = np.stack([synth_descriptors(MATCH_N, 100 * rank + i, base=base) for i in range(MATCH_SETS)])
    d_sets = torch.from_numpy(sets).to(dev)
    d_cnt = torch.full((MATCH_SETS,), MATCH_N, dtype=torch.int32, device=dev)
    d_pairs = torch.from_numpy(PL.all_pairs(MATCH_SETS)).to(dev)
    n_mpairs = int(d_pairs.shape[0])
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
    dpairs = float(n_mpairs) * MATCH_N * MATCH_N
    tc = mstat.get("k_match_tc", (0, 0.0))
    tc_ms = tc[1] / msteps if tc[1] else float("nan")
    tc_ach = 256.0 * dpairs / (tc_ms * 1e-3) / 1e12 if tc[1] else None
    match = {"metric": "nn_ratio_descriptor_pairs_per_s", "value": world * dpairs / (m_ms * 1e-3), "unit": "descriptor-pairs/s",
             "image_pairs_per_s": world * n_mpairs / (m_ms * 1e-3), "ms_per_step": m_ms, "steps": msteps,
             "workload": f"all {n_mpairs} pairs of {MATCH_SETS} sets of {MATCH_N} x 128 f32 descriptors per GPU (configs[4] pair shape), thr {RATIO}",
             "matches_per_pair_mean": float(mcnt.float().mean().item()),
             "rows_rescanned_frac": float(mst[:, 0].float().sum().item()) / (n_mpairs * MATCH_N),
             "candidate_groups_rechecked_per_row": float(mst[:, 1].float().sum().item()) / (n_mpairs * MATCH_N),
             "whole_path": {"achieved": 256.0 * dpairs / (m_ms * 1e-3) / 1e12, "unit": "TFLOP/s (algorithmic, every matcher kernel)",
                            "frac_of_sustained_peak": 256.0 * dpairs / (m_ms * 1e-3) / 1e12 / pk["tf_sust"],
                            "frac_of_burst_peak": 256.0 * dpairs / (m_ms * 1e-3) / 1e12 / pk["tf_burst"]},
             "kernels": {k: {"launches_per_step": v[0] / msteps, "ms_per_step": v[1] / msteps} for k, v in sorted(mstat.items())},
             "roofline": {"kernel": "k_match_tc (tcgen05 M128xN256xK16 fp16 GEMM, fused key + top-4 groups-of-8 epilogue)", "bound": "tensor",
                          "achieved": tc_ach, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                          "frac": (tc_ach / pk["tf_burst"]) if tc_ach else None,
                          "traffic": traffic.get("k_match_tc_dram_bytes_per_launch"),
                          "peak_sustained": pk["tf_sust"], "frac_of_sustained": (tc_ach / pk["tf_sust"]) if tc_ach else None,
                          "peak_source": pk["src"] + ", burst bf16 (the kernel is timed alone; fp16 runs at the same rate)",
                          "flops": "256 per descriptor pair (algorithmic == executed: single fp16 pass)"}}
    del d_sets, mm, mc

    # ---- configs[4] as written (every rank takes part)
    ap_leg = None
    if not args.no_all_pairs:
        ap_leg = all_pairs_leg(torch, dist, PL, rank, world, local, dev, pk, reps=1 if world == 1 else 2)

    # ---- configs[1] literally: ONE 1080p image (latency-bound: 47 MB of traffic, ~20 launches)
    single = None
    if rank == 0:
        from sfmfromscratch_b200 import ScaleRotInvSIFT
        one = images[:1]
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def ev_timed(fn, reps=20, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ev_a.record()
            for _ in range(reps):
                fn()
            ev_b.record()
            torch.cuda.synchronize()
            return ev_a.elapsed_time(ev_b) / reps
        lat = ev_timed(lambda: extract_batch_device(one, params, want_aux=False))
        ge = PL.GraphedExtractor(one, {})
        lat_graph = ev_timed(lambda: ge.run(one))
        gb = PL.GraphedExtractor(images[:batch], {})
        batch_graph = ev_timed(lambda: gb.run(images[:batch]))
        del ge, gb
        img_np = host_shard[0].numpy()
        ScaleRotInvSIFT(img_np, {})
        t0 = time.time()
        for _ in range(5):
            e = ScaleRotInvSIFT(img_np, {})
        host_ms = (time.time() - t0) / 5 * 1e3
        single = {"workload": "one 1920x1080 image, ScaleRotInvSIFT defaults (configs[1])",
                  "resident_ms": lat, "resident_mpixel_per_s": IMG_H * IMG_W / (lat * 1e-3) / 1e6,
                  "resident_cuda_graph_ms": lat_graph,
                  f"batch{batch}_extract_only_cuda_graph_ms": batch_graph,
                  "class_call_ms": host_ms, "class_call_mpixel_per_s": IMG_H * IMG_W / (host_ms * 1e-3) / 1e6,
                  "keypoints": int(len(e.detect_keypoints()[0])),
                  "algorithmic_bytes": 16.0 * lp + 4.0 * (lp - IMG_H * IMG_W),
                  "hbm_frac_resident": (16.0 * lp + 4.0 * (lp - IMG_H * IMG_W)) / (lat * 1e-3) / 1e9 / pk["hbm"],
                  "note": "class_call = ScaleRotInvSIFT(numpy image, {}) -> numpy keypoints/descriptors, pageable host memory in, "
                          "pinned staging cached per thread"}

        # ---- configs[0]: the reference's two-view chain on a 640x480 pair (Runner.py:49-63,347-351)
        from sfmfromscratch_b200 import NNRatioFeatureMatcher
        from sfmfromscratch_b200 import geometry as GE
        from sfmfromscratch_b200.synth import second_view
        a0 = synth_image(480, 640, 0)
        b0v = second_view(a0, 1)
        it = GE.calculate_num_ransac_iterations(0.98, 8, 0.4)

        def two_view_host():
            e1, e2 = ScaleRotInvSIFT(a0, {}), ScaleRotInvSIFT(b0v, {})
            (x1, y1), (x2, y2) = e1.detect_keypoints(), e2.detect_keypoints()
            m, c = NNRatioFeatureMatcher(RATIO).match_features_ratio_test(e1.extract_descriptors(), e2.extract_descriptors())
            p1, p2 = GE.convert_matches_to_coords(m, x1, y1, x2, y2, 2500)
            i1, i2 = GE.CameraPose.find_inliers(p1, p2, max_iterations=it)
            return len(x1), len(x2), len(m), len(i1)
        t0 = time.perf_counter()
        sizes0 = two_view_host()
        first_ms = (time.perf_counter() - t0) * 1e3
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            two_view_host()
            ts.append((time.perf_counter() - t0) * 1e3)
        d01 = torch.from_numpy(np.stack([a0, b0v])).to(dev)
        p01 = torch.tensor([[0, 1]], dtype=torch.int32, device=dev)
        smp = GE._samples_host(sizes0[2], it, GE.RANSAC_SEED).to(dev)

        def two_view_resident():
            o = extract_batch_device(d01, params, want_aux=False, check=False)
            m, c, cnt = match_batch_device(o['desc'], o['count'], p01, RATIO, cap=cap)
            q1, q2, _ = GE.matches_to_coords_device(m[0], cnt, o['x'][0], o['y'][0], o['x'][1], o['y'][1], 2500)
            GE.ransac_device(q1[:sizes0[2]], q2[:sizes0[2]], it, samples=smp)
        res0_ms = ev_timed(two_view_resident, reps=20, warm=3)
        config0 = {"workload": "configs[0]: two 640x480 views: ScaleRotInvSIFT x2 + NNRatioFeatureMatcher + _convert_matches_to_coords + "
                               f"CameraPose.find_inliers ({it} hypotheses), the reference's FeatureRunner / SFMRunner chain (Runner.py:49-63,347-351)",
                   "host_call_ms_median": float(np.median(ts)), "host_call_ms_min": float(np.min(ts)), "host_call_first_ms": first_ms,
                   "resident_ms": res0_ms, "keypoints": [sizes0[0], sizes0[1]], "matches": sizes0[2], "inliers": sizes0[3],
                   "note": "host_call: numpy images in, numpy inliers out, wall clock (the first call also draws the 8-subset table of "
                           "this match count); resident: images and the subset table in HBM, CUDA events, no host read in the chain"}
    else:
        config0 = None

    # ---- configs[2]: two 3840x2160 views at ~20 k keypoints, matched; device-resident, CUDA events, no host read
    cfg2 = None
    if rank == 0 and not args.no_4k:
        from sfmfromscratch_b200.synth import second_view
        a4 = synth_image(2160, 3840, 5)
        b4 = second_view(a4, 6)
        d4 = torch.from_numpy(np.stack([a4, b4])).to(dev)
        p4, keep4 = make_params({'num_interest_points': 32000}, pyramid=True)
        pr4 = torch.tensor([[0, 1]], dtype=torch.int32, device=dev)
        res4 = {}

        def run4():
            o = extract_batch_device(d4, p4, want_aux=False, check=False)
            res4['o'] = o
            res4['m'] = match_batch_device(o['desc'], o['count'], pr4, RATIO, cap=32000)   # counts stay on the device
        N.profile_enable(True, local)
        ms4 = ev_timed(run4, reps=10, warm=2)
        st4 = N.profile_collect(local)
        N.profile_enable(False, local)
        n4 = [int(v) for v in res4['o']['count'].cpu()]
        ext4 = sum(v[1] for k, v in st4.items() if not k.startswith("k_match") and k != "k_work_scan") / 12
        cfg2 = {"workload": "configs[2]: two 3840x2160 views, num_interest_points 32000: extraction of both + NN-ratio matching",
                "ms": ms4, "extract_ms": ext4, "match_ms": ms4 - ext4, "keypoints": n4, "matches": int(res4['m'][2].cpu()[0]),
                "mpixel_per_s": 2 * 2160 * 3840 / (ms4 * 1e-3) / 1e6,
                "descriptor_pairs_per_s": float(n4[0]) * n4[1] / ((ms4 - ext4) * 1e-3) if ms4 > ext4 else None,
                "note": "CUDA events, images resident, keypoint counts passed to the matcher on the device (no host read between the stages)"}
        del keep4, d4, res4

    # ---- two-view RANSAC and point association on the matcher's output
    geom = None
    if rank == 0 and not args.no_geometry:
        geom = geometry_leg(torch, N, local, cpu=(world == 1 and not args.no_cpu))

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        from oracle import stage_reference
        n_cpu = 8
        frames = cpu_frames(n_cpu)
        v, dt = cpu_sample(frames, _port_extract, _port_match)
        cpu = {"value": v, "unit": "Mpixel/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"{n_cpu} of the job's {N_IMAGES} 1080p frames + their {n_cpu - 1} consecutive-pair matches through the "
                         f"CPU oracle (C + numpy port of the reference), one process"}
        ref_dir = stage_reference.staged()
        if ref_dir and not args.no_ref_sample:
            # the reference itself on ONE frame pair (its Python NMS loop needs ~15 s per 1080p frame)
            import subprocess
            code = ("import sys, json, time; sys.path.insert(0, %r); import bench; bench._ref_init(%r); "
                    "fr = bench.cpu_frames(2); t = time.time(); f = [bench._ref_extract(x) for x in fr]; t1 = time.time(); "
                    "k = bench._ref_match((f[0], f[1])); t2 = time.time(); "
                    "print(json.dumps({'extract_s': t1 - t, 'match_s': t2 - t1, 'keypoints': [len(f[0]), len(f[1])], 'matches': k}))"
                    % (ROOT, ref_dir))
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
                d = json.loads(r.stdout.strip().splitlines()[-1])
                tot = d['extract_s'] + d['match_s']
                cpu["reference_python"] = {"value": 2 * IMG_H * IMG_W / tot / 1e6, "unit": "Mpixel/s", "cores": 1, "kind": "reference",
                                           "seconds": tot, **d,
                                           "sample": "2 of the job's frames + their pair through the UNMODIFIED reference classes "
                                                     "(oracle/_ref), one process"}
            except Exception as ex:                       # the baseline is a reported extra, never a reason to lose the line
                cpu["reference_python"] = {"unavailable": repr(ex)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "images": n_job, "images_per_gpu": per, "batch": batch, "image": [IMG_H, IMG_W],
                           "distinct_images": n_job, "pairs": int(len(pairs_global)), "pairs_per_gpu": n_my_pairs,
                           "keypoints_per_image": kp_per_image, "matches_per_pair": matches_per_pair,
                           "sequence": "frames pan over one blurred-noise canvas by 24 px each, per-frame N(0, 0.005) noise "
                                       "(sfmfromscratch_b200.synth.frame_sequence)",
                           "l2": f"inputs larger than L2 ({per * IMG_H * IMG_W * 4 / 1e6:.0f} MB of distinct frames per GPU and step, "
                                 f"{batch * lp * 4 / 1e6:.0f} MB of R planes per batch)",
                           "step": ("FeaturePipeline.stream_resident: extraction on the current stream, exchange + matching of the same job on a "
                                    "second stream under the next job's extraction (ms_per_step_serial = everything on one stream; the "
                                    "per-kernel times come from that serial form)") if not args.serial_step else "extraction, exchange, matching on one stream",
                           "parallelism": f"image shards x{world}, all-gather (NCCL) of the {plan.K} descriptor block(s) per rank that "
                                          f"another shard's pairs need (policy {plan.policy}), pair shards"},
                "roofline": roofline, "roofline_nms": roofline_nms, "extraction": extraction,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "sustained": sustained, "exchange_ms": exchange_ms,
                "ms_per_step_serial": serial_ms, "ms_per_step_profiled": prof_ms, "kernels": kernels, "all_pairs": ap_leg, "match": match,
                "single_image": single, "config0_two_view": config0, "config2_4k_pair": cfg2, "geometry": geom}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    del keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-sample", action="store_true", help="skip the one-pair run of the staged reference inside cpu_baseline")
    ap.add_argument("--no-4k", action="store_true", help="skip the configs[2] leg")
    ap.add_argument("--no-all-pairs", action="store_true", help="skip the configs[4] leg")
    ap.add_argument("--no-geometry", action="store_true", help="skip the RANSAC / association leg (SURVEY 8f rows 2-3)")
    ap.add_argument("--port-only", action="store_true", help="--impl reference: time the oracle port only")
    ap.add_argument("--serial-step", action="store_true", help="matching behind the extraction on one stream (no overlap across steps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
