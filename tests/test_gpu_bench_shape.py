"""Parity at the shapes bench.py measures (BASELINE.json configs[2], [3], [4]): the batched 32 x 1080p
extraction with interior TMA tiles over the batch dimension, fused next-level emission and the deferred
overflow check; consecutive pairs through the pair plan; 4K descriptors; matcher batches above the
65 535-pair grid limit.  Needs a B200: pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from parity import (DescriptorExplainer, assert_descriptors_close, assert_keypoints_equal,  # noqa: E402
                    assert_matches_identical)


@pytest.fixture(scope="module")
def bench_batch():
    """One bench batch: 32 distinct 1080p frames of the bench sequence, extracted exactly as bench.py's step
    does (FeaturePipeline.extract into slices of shard-wide tensors, deferred overflow check)."""
    import torch
    from sfmfromscratch_b200 import pipeline as PL
    from sfmfromscratch_b200.synth import frame_sequence
    frames = frame_sequence(1080, 1920, 0, 32, 256, threads=8)
    dev = torch.device("cuda", 0)
    images = torch.from_numpy(frames).to(dev)
    pipe = PL.FeaturePipeline({}, 0.8)
    cap = 2500
    i32 = dict(dtype=torch.int32, device=dev)
    shard = {'x': torch.empty((32, cap), **i32), 'y': torch.empty((32, cap), **i32), 'count': torch.empty((32,), **i32),
             'desc': torch.empty((32, cap, 128), dtype=torch.float32, device=dev)}
    pipe.extract(images, deferred_check=True, out=shard)
    assert not pipe.overflow_since_last_check()
    plan = pipe.pair_plan(PL.consecutive_pairs(32), 32)
    m = pipe.match_plan(plan, shard['desc'], shard['count'], cap=cap)
    torch.cuda.synchronize()
    host = {k: v.cpu().numpy() for k, v in shard.items()}
    return frames, host, plan, tuple(t.cpu().numpy() for t in m)


def test_batch32_1080p_equals_single_image_calls(bench_batch):
    """Every image of the 32-frame batch returns exactly what a single-image class call returns."""
    import sfmfromscratch_b200 as S
    frames, host, _, _ = bench_batch
    for b in range(32):
        e = S.ScaleRotInvSIFT(frames[b], {})
        X, Y = e.detect_keypoints()
        n = int(host['count'][b])
        assert n == len(X)
        assert np.array_equal(host['x'][b, :n], X) and np.array_equal(host['y'][b, :n], Y)
        assert np.array_equal(host['desc'][b, :n], e.extract_descriptors())


@pytest.mark.parametrize("b", [0, 11, 22, 31])
def test_batch32_1080p_equals_oracle(bench_batch, b):
    """Four of the batch's frames against the CPU oracle: keypoints identical, descriptors within tolerance
    with no keypoint needing the bin-edge exemption."""
    from oracle import oracle as O
    frames, host, _, _ = bench_batch
    o = O.ScaleRotInvSIFT(frames[b], {})
    n = int(host['count'][b])
    assert_keypoints_equal(host['x'][b, :n].astype(np.int64), host['y'][b, :n].astype(np.int64), *o.detect_keypoints())
    ex = DescriptorExplainer(frames[b], {}, o.levels, o.level_x, o.level_y)
    assert assert_descriptors_close(host['desc'][b, :n], o.extract_descriptors(), ex) == 0


def test_consecutive_pairs_through_pair_plan_equal_oracle(bench_batch):
    """The step's matcher call (31 consecutive pairs through PairPlan / match_plan) against the oracle matcher
    on the same descriptors: (row, index, confidence) bit-identical for every pair."""
    from oracle import oracle as O
    _, host, plan, (mm, mc, mn) = bench_batch
    assert len(plan.mine) == 31
    total = 0
    for k, (i, j) in enumerate(plan.mine.tolist()):
        ni, nj = int(host['count'][i]), int(host['count'][j])
        mo, co = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(host['desc'][i, :ni], host['desc'][j, :nj])
        n = int(mn[k])
        assert n == len(mo)
        assert_matches_identical(mm[k, :n], mc[k, :n], mo, co)
        total += n
    assert total > 31 * 500          # consecutive frames overlap: true correspondences exist


def test_4k_descriptors_vs_oracle():
    """configs[2] extraction: 3840x2160 at num_interest_points 32000 (~19 k keypoints): keypoints AND descriptors
    against the oracle."""
    from oracle import oracle as O
    import sfmfromscratch_b200 as S
    from sfmfromscratch_b200.synth import synth_image
    img = synth_image(2160, 3840, 1)
    p = {'num_interest_points': 32000}
    g = S.ScaleRotInvSIFT(img, p)
    o = O.ScaleRotInvSIFT(img, p)
    X, Y = g.detect_keypoints()
    assert len(X) > 15000
    assert_keypoints_equal(X, Y, *o.detect_keypoints())
    assert np.array_equal(g.confidences.view(np.uint32), o.confidences.view(np.uint32))
    ex = DescriptorExplainer(img, p, o.levels, o.level_x, o.level_y)
    licensed = assert_descriptors_close(g.extract_descriptors(), o.extract_descriptors(), ex)
    assert licensed <= 2             # a 4 ulp window around 45 edges x 5 M samples: a hit is possible, a handful is not


def test_matcher_batch_above_grid_limit_and_bad_pair_ids():
    """More than 65 535 pairs in one call (gridDim.y of the per-row kernels): the library runs them as chunks.
    Pair ids outside the set table are matched as empty sets (count 0), nothing is read out of bounds."""
    import torch
    from sfmfromscratch_b200.matcher import match_batch_device
    rng = np.random.default_rng(5)
    n_sets, nmax = 40, 24
    desc = np.sqrt(rng.gamma(0.5, 1.0, size=(n_sets, nmax, 128))).astype(np.float32)
    desc[:, ::3] = desc[0, ::3]                       # shared rows: true matches between any two sets
    desc += rng.normal(0, 0.01, desc.shape).astype(np.float32)
    counts = rng.integers(2, nmax + 1, n_sets).astype(np.int32)
    P = 70000
    pairs = rng.integers(0, n_sets, (P, 2)).astype(np.int32)
    bad = rng.choice(P, 50, replace=False)
    pairs[bad[:25], 0] = n_sets + 3
    pairs[bad[25:], 1] = -1
    dev = torch.device("cuda", 0)
    m, c, cnt = match_batch_device(torch.from_numpy(desc).to(dev), torch.from_numpy(counts).to(dev),
                                   torch.from_numpy(pairs).to(dev), 0.8, cap=nmax)
    m, c, cnt = m.cpu().numpy(), c.cpu().numpy(), cnt.cpu().numpy()
    assert (cnt[bad] == 0).all()
    from oracle import oracle as O
    check = np.concatenate([np.arange(0, 200), np.arange(65400, 65700), np.arange(P - 200, P)])
    seen = 0
    for k in check:
        if k in bad:
            continue
        i, j = pairs[k]
        mo, co = O.NNRatioFeatureMatcher(0.8).match_features_ratio_test(desc[i, :counts[i]], desc[j, :counts[j]])
        assert cnt[k] == len(mo)
        assert_matches_identical(m[k, :cnt[k]], c[k, :cnt[k]], mo, co)
        seen += len(mo)
    assert seen > 0
    # the whole batch against a second call over a different chunking of the same pairs
    m2, c2, cnt2 = match_batch_device(torch.from_numpy(desc).to(dev), torch.from_numpy(counts).to(dev),
                                      torch.from_numpy(pairs[30000:]).to(dev), 0.8, cap=nmax)
    assert np.array_equal(cnt[30000:], cnt2.cpu().numpy())


def test_stream_resident_equals_serial_step(bench_batch):
    """bench.py's step (FeaturePipeline.stream_resident: matching on a second stream under the next job's extraction,
    two table sets used alternately) returns, job after job, exactly what extraction followed by match_plan returns."""
    import torch
    from sfmfromscratch_b200 import pipeline as PL
    frames, host, plan, (mm, mc, mn) = bench_batch
    dev = torch.device("cuda", 0)
    images = torch.from_numpy(frames).to(dev)
    pipe = PL.FeaturePipeline({}, 0.8)
    plan2 = pipe.pair_plan(PL.consecutive_pairs(32), 32)
    jobs = []
    for _ in range(4):                                   # both table sets twice, no host wait in between
        tabs, m, ev = pipe.stream_resident(images, plan2, batch=8, cap=2500)
        jobs.append((tabs, m, ev))
    pipe.join_resident()
    torch.cuda.synchronize()
    assert not pipe.overflow_since_last_check()
    assert jobs[0][0]['desc'].data_ptr() == jobs[2][0]['desc'].data_ptr() != jobs[1][0]['desc'].data_ptr()
    for tabs, m, ev in jobs[2:]:                         # (the first two jobs' tables have been overwritten by these)
        assert ev.query()
        for k in ('x', 'y', 'count', 'desc'):
            got = tabs[k].cpu().numpy()
            if k == 'count':
                assert np.array_equal(got[:32], host[k])
            else:
                for b in range(32):
                    n = int(host['count'][b])
                    assert np.array_equal(got[b, :n], host[k][b, :n])
        for a, b in zip(m, (mm, mc, mn)):
            a = a.cpu().numpy()
            for k in range(31):
                n = int(mn[k])
                assert np.array_equal(a[k][:n] if a.ndim > 1 else a[k], b[k][:n] if b.ndim > 1 else b[k])
