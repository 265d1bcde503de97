"""Batch / multi-GPU orchestration of the hot path.

One process per GPU.  Images are sharded in contiguous blocks for extraction,
descriptor blocks are exchanged with ONE collective (all-gather over NCCL /
NVLink; gloo on CPU in the tests), and image pairs are dealt block-cyclically
for matching.  Nothing else communicates: both stages are embarrassingly
parallel, so no other collective is invented.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_images(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of ceil(n / world) images for `rank`."""
    per = (n_images + world - 1) // world
    start = min(rank * per, n_images)
    return start, min(start + per, n_images)


def consecutive_pairs(n_images: int) -> np.ndarray:
    """(i, i+1) for every consecutive image pair -- what Runner.py:183-191 matches."""
    i = np.arange(max(n_images - 1, 0), dtype=np.int32)
    return np.stack([i, i + 1], axis=1)


def all_pairs(n_images: int) -> np.ndarray:
    """Unordered pairs i < j in row-major order (configs[4]: 512 images -> 130 816 pairs)."""
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)


def deal_pairs(pairs: np.ndarray, rank: int, world: int, block: int = 64) -> np.ndarray:
    """Block-cyclic deal: blocks of `block` consecutive pairs go to ranks in turn,
    so every rank sees the same mix of images (and of set sizes)."""
    n = len(pairs)
    idx = np.arange(n)
    mine = (idx // block) % world == rank
    return pairs[mine]


def gather_descriptors(desc: torch.Tensor, counts: torch.Tensor, group=None):
    """All-gather the per-rank descriptor blocks [b, nmax, 128] (+ counts [b])
    into [world*b, nmax, 128] / [world*b] on every rank.  Every rank must
    contribute the same b and nmax (pad the last shard)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return desc, counts
    world = dist.get_world_size(group)
    desc = desc.contiguous()
    counts = counts.contiguous()
    out_d = torch.empty((world * desc.shape[0],) + tuple(desc.shape[1:]), dtype=desc.dtype, device=desc.device)
    out_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    try:
        dist.all_gather_into_tensor(out_d, desc, group=group)
        dist.all_gather_into_tensor(out_c, counts, group=group)
    except (RuntimeError, NotImplementedError):          # backends without the fused form
        dl = list(out_d.chunk(world, dim=0))
        cl = list(out_c.chunk(world, dim=0))
        dist.all_gather(dl, desc, group=group)
        dist.all_gather(cl, counts, group=group)
    return out_d, out_c


def gather_keypoints(x: torch.Tensor, y: torch.Tensor, group=None):
    """All-gather the per-rank keypoint coordinates [b, cap] (int32) into [world*b, cap] on every
    rank: the geometry stage of a pair needs the coordinates of both images (Runner.py:347)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x, y
    world = dist.get_world_size(group)
    xy = torch.stack([x, y], dim=0).contiguous()                       # [2, b, cap]
    out = torch.empty((world,) + tuple(xy.shape), dtype=xy.dtype, device=xy.device)
    try:
        dist.all_gather_into_tensor(out, xy, group=group)
    except (RuntimeError, NotImplementedError):
        dist.all_gather(list(out.unbind(0)), xy, group=group)
    b = x.shape[0]
    return (out[:, 0].reshape(world * b, -1), out[:, 1].reshape(world * b, -1))


class PairPlan:
    """Who matches which pair, and which descriptor blocks have to travel for it.  Every rank builds the
    same plan from the global pair list (image ids over `world` contiguous shards of `per` images).

    * policy "local": a pair belongs to the rank that owns its first image.  Chosen when that is
      balanced and few blocks are needed remotely -- the reference's own pattern, consecutive pairs
      (Runner.py:183-191), needs ONE block per rank (the next shard's first image), so the exchange is
      an all-gather of 1.3 MB per rank instead of every descriptor of every rank.
    * policy "all": block-cyclic deal and the full all-gather (all-pairs matching, configs[4]:
      every block is needed everywhere).

    The matcher sees a local table: this rank's `per` blocks followed by the gathered blocks
    (`world * K` for "local", K = the largest number of blocks any rank has to send); `pairs_local`
    indexes that table, `mine` keeps the global ids in the same order."""

    def __init__(self, pairs_global: np.ndarray, per: int, rank: int, world: int, block: int = 64):
        pairs_global = np.asarray(pairs_global, dtype=np.int32).reshape(-1, 2)
        self.per, self.rank, self.world = per, rank, world
        owner = pairs_global // max(per, 1)                       # [P, 2] owning rank of each image
        first = owner[:, 0]
        per_rank = np.bincount(first, minlength=world)[:world] if len(first) else np.zeros(world, dtype=np.int64)
        send = [np.zeros(0, dtype=np.int32) for _ in range(world)]
        if world > 1 and len(pairs_global):
            remote = owner[:, 0] != owner[:, 1]                   # second image lives elsewhere
            ids = np.unique(pairs_global[remote, 1])
            for r in range(world):
                send[r] = ids[(ids // per) == r].astype(np.int32)
        K = max((len(x) for x in send), default=0)
        mean = max(len(pairs_global) / max(world, 1), 1.0)
        balanced = len(pairs_global) == 0 or per_rank.max() <= 1.25 * mean + 1
        if world == 1 or (balanced and K * world <= 0.5 * per * world):
            self.policy, self.K, self.send = "local", (K if world > 1 else 0), send
            self.mine = pairs_global[first == rank] if world > 1 else pairs_global
            loc = np.empty_like(self.mine)
            loc[:, 0] = self.mine[:, 0] - (rank * per if world > 1 else 0)
            j = self.mine[:, 1].astype(np.int64)
            if world == 1:
                loc[:, 1] = j
            else:
                own = (j // per) == rank
                # slot of a remote image: after this rank's `per` blocks, rank q's K gathered blocks start at
                # per + q * K, in the order of send[q] (sorted ids: position by binary search)
                sent = np.concatenate(send) if K else np.zeros(0, dtype=np.int32)
                first_of = np.cumsum([0] + [len(x) for x in send])[:-1]
                q = np.minimum(j // per, world - 1)
                within = np.searchsorted(sent, j) - first_of[q] if K else np.zeros_like(j)
                loc[:, 1] = np.where(own, j - rank * per, per + q * self.K + within)
            self.pairs_local = loc
            self.send_local = (send[rank] - rank * per).astype(np.int64) if world > 1 else np.zeros(0, dtype=np.int64)
        else:
            self.policy, self.K, self.send = "all", per, send
            self.mine = deal_pairs(pairs_global, rank, world, block=block)
            self.pairs_local = self.mine                          # global ids index the gathered table directly
            self.send_local = np.arange(per, dtype=np.int64)
        self._dev = {}

    def pairs_dev(self, device) -> Optional[torch.Tensor]:
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(np.ascontiguousarray(self.pairs_local)).to(device)
                              if len(self.pairs_local) else None)
        return self._dev[key]

    def send_idx_dev(self, device) -> torch.Tensor:
        """Local indices of the blocks this rank contributes, on `device` (uploaded once)."""
        key = "send:" + str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.send_local).to(device)
        return self._dev[key]


def exchange_for(plan: PairPlan, desc: torch.Tensor, counts: torch.Tensor, group=None):
    """The descriptor table `plan.pairs_local` indexes: everything (policy "all": one all-gather of the
    whole blocks) or this rank's blocks followed by the few blocks other ranks contribute (policy
    "local": one all-gather of K blocks per rank; no communication at all when K == 0)."""
    if plan.policy == "all":
        return gather_descriptors(desc, counts, group)
    if plan.K == 0:
        return desc, counts
    import torch.distributed as dist
    idx = plan.send_idx_dev(desc.device)
    world = dist.get_world_size(group)
    K, n_own = plan.K, len(plan.send_local)
    own_d, own_c = desc[:plan.per], counts[:plan.per]           # (the table may carry the exchange slots behind them)
    blk = int(np.prod(desc.shape[1:]))                         # floats per descriptor block
    # ONE collective: the K blocks and their K counts travel in one packed float32 buffer (the counts as raw int32
    # bits behind the blocks); two small all-gathers cost their launch latency twice (~0.05 ms per step at N = 8).
    # The two buffers live with the plan (allocated and zeroed once per device and block shape).
    kpad = (K + 3) // 4 * 4
    key = ("xbuf", str(desc.device), blk, world)
    if key not in plan._dev:
        plan._dev[key] = (torch.zeros((K * blk + kpad,), dtype=torch.float32, device=desc.device),
                          torch.empty((world, K * blk + kpad), dtype=torch.float32, device=desc.device))
    send, got = plan._dev[key]
    if n_own:
        torch.index_select(own_d.reshape(plan.per, blk), 0, idx, out=send[:n_own * blk].view(n_own, blk))
        send[K * blk:K * blk + n_own].view(torch.int32).copy_(own_c.index_select(0, idx))
    try:
        dist.all_gather_into_tensor(got, send, group=group)
    except (RuntimeError, NotImplementedError):                # backends without the fused form
        dist.all_gather(list(got.unbind(0)), send, group=group)
    got_d = got[:, :K * blk].reshape(world * K, *desc.shape[1:])            # strided view
    got_c = got[:, K * blk:K * blk + K].view(torch.int32).reshape(world * K).to(counts.dtype)
    if desc.shape[0] == plan.per + world * K:
        # the caller's table already has the slots behind this rank's blocks (FeaturePipeline.shard_tables): only the
        # gathered blocks move (world * K blocks instead of a copy of the whole shard, 164 MB per step at N = 2)
        desc[plan.per:].copy_(got_d)
        counts[plan.per:].copy_(got_c)
        return desc, counts
    return torch.cat([desc, got_d], dim=0), torch.cat([counts, got_c], dim=0)


class GraphedExtractor:
    """sfm_extract_batch for a fixed batch shape captured once into a CUDA graph: the ~25 kernel
    launches of an extraction replay as one submission (what matters for a single image, where
    launch latency, not bandwidth, sets the time).  `sample` is a representative float32 CUDA
    batch [B, H, W]; run(images) copies into the static input and replays; results are the
    static output tensors (valid until the next run)."""

    def __init__(self, sample: torch.Tensor, extractor_params: Optional[dict] = None, pyramid: bool = True):
        from .extractor import check_extract_status, extract_batch_device, make_params
        self.params, self._keep = make_params(extractor_params, pyramid=pyramid)
        self.static_in = sample.clone()
        extract_batch_device(self.static_in, self.params, want_aux=False)          # warm-up; sizes candidate buffers
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = extract_batch_device(self.static_in, self.params, want_aux=False, check=False)
        self._check = check_extract_status

    def run(self, images: torch.Tensor, check: bool = False):
        self.static_in.copy_(images, non_blocking=True)
        self.graph.replay()
        if check and not self._check(self.out):
            raise RuntimeError("candidate buffer overflow in a graphed extraction: rebuild with cand_full")
        return self.out


class FeaturePipeline:
    """Extraction of this rank's image shard, descriptor exchange, matching of
    this rank's pair share.  All tensors stay on the device."""

    def __init__(self, extractor_params: Optional[dict] = None, ratio_threshold: float = 0.8,
                 rank: int = 0, world: int = 1, group=None):
        from .extractor import make_params
        self.params, self._keep = make_params(extractor_params, pyramid=True)
        self.ratio_threshold = ratio_threshold
        self.rank, self.world, self.group = rank, world, group
        self.pair_block = 64

    def extract(self, images: torch.Tensor, deferred_check: bool = False, out: Optional[dict] = None):
        """Extraction of a resident batch.  By default the candidate-overflow flag is read back right
        away (a host wait per call).  With `deferred_check` the flags accumulate on the device -- nothing
        waits, the host keeps enqueueing -- and `overflow_since_last_check()` reads them once.  `out` may hold
        preallocated result tensors (x, y, count, desc), e.g. slices of a shard-wide table filled batch by batch."""
        from .extractor import extract_batch_device
        out = extract_batch_device(images, self.params, want_aux=False, check=not deferred_check, out=out)
        if deferred_check:
            if getattr(self, '_flag_acc', None) is None or self._flag_acc.device != images.device:
                self._flag_acc = torch.zeros((1,), dtype=torch.int32, device=images.device)
            self._flag_acc |= out['_flag'][:1]
        return out

    def overflow_since_last_check(self) -> bool:
        """True when some `extract(deferred_check=True)` since the last call hit a plateau image: those
        batches must be redone with `params.cand_full = 1` (what the immediate check does by itself)."""
        acc = getattr(self, '_flag_acc', None)
        if acc is None:
            return False
        bad = int(acc.cpu()[0]) != 0
        acc.zero_()
        return bad

    def exchange(self, desc: torch.Tensor, counts: torch.Tensor):
        return gather_descriptors(desc, counts, self.group)

    def shard_tables(self, plan: PairPlan, cap: int, device) -> dict:
        """Result tables of this rank's `plan.per` images (x, y, count, desc) whose descriptor / count tables carry, behind
        the rank's own blocks, the slots the exchange of `plan` fills (policy "local": world * K blocks) -- extraction
        writes rows [0, per), `match_plan` gathers in place and nothing is concatenated per step."""
        extra = plan.world * plan.K if (plan.policy == "local" and plan.world > 1) else 0
        i32 = dict(dtype=torch.int32, device=device)
        return {'x': torch.empty((plan.per, cap), **i32), 'y': torch.empty((plan.per, cap), **i32),
                'count': torch.zeros((plan.per + extra,), **i32),
                'desc': torch.empty((plan.per + extra, cap, 128), dtype=torch.float32, device=device)}

    def pair_plan(self, pairs_global: np.ndarray, per: int) -> PairPlan:
        """The (cached) PairPlan of `pairs_global` for shards of `per` images."""
        key = (np.asarray(pairs_global).tobytes(), per, self.rank, self.world, self.pair_block)
        if getattr(self, '_plan_key', None) != key:
            self._plan_key, self._plan = key, PairPlan(pairs_global, per, self.rank, self.world, block=self.pair_block)
        return self._plan

    def match_plan(self, plan: PairPlan, desc: torch.Tensor, counts: torch.Tensor, cap: Optional[int] = None):
        """Exchange what `plan` needs and match this rank's pairs; None when it has none."""
        tab_d, tab_c = exchange_for(plan, desc, counts, self.group)
        if not len(plan.mine):
            return None
        if plan.policy == "all":
            return self.match(tab_d, tab_c, plan.pairs_dev(desc.device), cap=cap, pairs_host=plan.mine)
        from .matcher import match_batch_device
        return match_batch_device(tab_d, tab_c, plan.pairs_dev(desc.device), self.ratio_threshold, cap=cap)

    def match(self, desc_all: torch.Tensor, counts_all: torch.Tensor, pairs: torch.Tensor, cap: Optional[int] = None,
              pairs_host: Optional[np.ndarray] = None):
        """Match `pairs` (ids into desc_all).  When this rank's pairs touch only a fraction of the
        gathered sets (pair sharding), the referenced blocks are compacted first so the matcher's
        per-set preparation does not run over every rank's descriptors."""
        from .matcher import match_batch_device
        if pairs.shape[0] == 0:
            return None
        if pairs_host is not None and desc_all.shape[0] > 1:
            ids, inv = np.unique(pairs_host.reshape(-1), return_inverse=True)
            if len(ids) <= 0.6 * desc_all.shape[0]:
                idx = torch.from_numpy(ids.astype(np.int64)).to(desc_all.device, non_blocking=True)
                desc_all = desc_all.index_select(0, idx)
                counts_all = counts_all.index_select(0, idx)
                pairs = torch.from_numpy(inv.reshape(-1, 2).astype(np.int32)).to(desc_all.device, non_blocking=True)
        return match_batch_device(desc_all, counts_all, pairs, self.ratio_threshold, cap=cap)

    def stream_resident(self, images: torch.Tensor, plan: PairPlan, batch: int = 32, cap: Optional[int] = None):
        """One job over a RESIDENT shard, for a sequence of jobs: extraction of `images` ([per, H, W] on the device) in
        calls of `batch` frames on the current stream, then the exchange `plan` needs and the matching of this rank's
        pairs on a second stream -- so job k's matcher (latency-bound gathers that leave most of the SMs' issue slots
        idle) runs under job k+1's extraction instead of behind it.  Two sets of result tables are used alternately; a
        set is reused only after the matching that read it has finished (events, no host wait).  Candidate-overflow
        flags accumulate as in `extract(deferred_check=True)`.  Returns (tables, match result or None, completion
        event): the tables and the match tensors are valid once the event has completed, and until the job after next."""
        dev = images.device
        main = torch.cuda.current_stream()
        per = images.shape[0]
        cap = cap or self.max_keypoints()
        key = (id(plan), per, cap, dev.index)
        if getattr(self, '_res_key', None) != key:
            if getattr(self, '_res_sets', None):
                for st in self._res_sets:
                    if st['free'] is not None:
                        st['free'].synchronize()
            self._s_match = getattr(self, '_s_match', None) or torch.cuda.Stream()
            self._res_sets = [dict(self.shard_tables(plan, cap, dev), free=None) for _ in range(2)]
            self._res_key, self._res_i = key, 0
        st = self._res_sets[self._res_i]
        self._res_i ^= 1
        if st['free'] is not None:
            main.wait_event(st['free'])                         # the matching that last read this set is done
        tabs = {k: st[k] for k in ('x', 'y', 'count', 'desc')}
        for b0 in range(0, per, batch):
            b1 = min(b0 + batch, per)
            self.extract(images[b0:b1], deferred_check=True, out={k: v[b0:b1] for k, v in tabs.items()})
        ready = torch.cuda.Event()
        ready.record(main)
        self._s_match.wait_event(ready)
        with torch.cuda.stream(self._s_match):
            m = self.match_plan(plan, tabs['desc'], tabs['count'], cap=cap)
            st['free'] = torch.cuda.Event()
            st['free'].record(self._s_match)
        return tabs, m, st['free']

    def join_resident(self) -> None:
        """The current stream waits for the matching `stream_resident` has put on its second stream (no host wait)."""
        if getattr(self, '_s_match', None) is not None:
            torch.cuda.current_stream().wait_stream(self._s_match)

    def max_keypoints(self) -> int:
        """Rows of the result tables: sfm_extract_max_keypoints of this pipeline's parameters."""
        import ctypes as C
        from . import _native as N
        return int(N.load_library().sfm_extract_max_keypoints(C.byref(self.params)))

    def run_host(self, host_images: torch.Tensor, pairs_global: np.ndarray, host_out: dict, chunk: int = 8,
                 _params=None):
        """The hot path end to end from HOST buffers: `host_images` is a pinned
        float32 [B, H, W] tensor, `host_out` holds pinned result tensors
        (x, y, desc, count [, matches, conf, mcount]).  The batch is cut into
        chunks; chunk i+1's host-to-device copy and chunk i-1's device-to-host
        copy run on their own streams while chunk i is being extracted, so the
        PCIe transfers overlap the kernels (both DMA directions are independent).
        (Measured and dropped: tapering the chunk sizes and matching pairs incrementally as their
        images finish -- with ~25 launches per chunk the host enqueue rate, not the tail, becomes the
        limit: 6.2-6.6 ms against 5.9 ms.)  Returns this rank's pair list, in the order of the rows of
        host_out['matches'], after everything has landed in `host_out`."""
        from .extractor import check_extract_status, copy_params, extract_batch_device
        params = self.params if _params is None else _params
        dev = torch.device('cuda', torch.cuda.current_device())
        main = torch.cuda.current_stream()
        if not hasattr(self, '_s_in'):
            self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
        B, H, W = host_images.shape
        cap = host_out['x'].shape[1]
        imgs = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        full = {'x': torch.empty((B, cap), **i32), 'y': torch.empty((B, cap), **i32),
                'count': torch.empty((B,), **i32),
                'desc': torch.empty((B, cap, 128), dtype=torch.float32, device=dev)}
        self._s_in.wait_stream(main)                 # the allocations above are stream-ordered on `main`
        self._s_out.wait_stream(main)
        bounds = [(c0, min(c0 + chunk, B)) for c0 in range(0, B, chunk)]
        ready = []
        with torch.cuda.stream(self._s_in):
            for c0, c1 in bounds:
                imgs[c0:c1].copy_(host_images[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._s_in)
                ready.append(ev)
        flags = []
        for (c0, c1), ev in zip(bounds, ready):
            main.wait_event(ev)
            res = extract_batch_device(imgs[c0:c1], params, want_aux=False, check=False,
                                       out={k: v[c0:c1] for k, v in full.items()})
            flags.append(res)
            done = torch.cuda.Event()
            done.record(main)
            self._s_out.wait_event(done)
            with torch.cuda.stream(self._s_out):
                for k in ('x', 'y', 'desc', 'count'):
                    host_out[k][c0:c1].copy_(full[k][c0:c1], non_blocking=True)
        plan = self.pair_plan(pairs_global, B)
        mine = plan.mine
        m = self.match_plan(plan, full['desc'], full['count'], cap=host_out['matches'].shape[1])
        if m is not None:
            host_out['matches'][:len(mine)].copy_(m[0], non_blocking=True)
            host_out['conf'][:len(mine)].copy_(m[1], non_blocking=True)
            host_out['mcount'][:len(mine)].copy_(m[2], non_blocking=True)
        main.wait_stream(self._s_out)
        for t in (imgs, *full.values()):
            t.record_stream(self._s_in); t.record_stream(self._s_out)
        main.synchronize()
        if not all(check_extract_status(r) for r in flags):
            # plateau image somewhere in the batch: redo it ONCE with full-size candidate buffers, on a private
            # copy of the parameters (self.params is shared with other threads and with later calls)
            if params.cand_full:
                raise RuntimeError("candidate buffer overflow with full-size buffers")
            retry = copy_params(params)
            retry.cand_full = 1
            return self.run_host(host_images, pairs_global, host_out, chunk, _params=retry)
        return mine

    def stream_host(self, host_images: torch.Tensor, pairs_global: np.ndarray, host_out: dict, chunk: int = 8):
        """`run_host` for a SEQUENCE of batches: nothing here waits on the host, so the host-to-device copy of
        batch k+1 runs under the matching / read-back tail of batch k (a lone `run_host` leaves the PCIe link
        idle for ~1 ms per step at both ends).  Device buffers are two persistent slots used alternately; a slot
        is recycled only after the kernels and copies of the batch that used it two calls ago have finished
        (events, no host sync).  Pass a different `host_out` to consecutive calls and read one only after
        `drain()` (or after the event this returns has completed).  Returns (pairs of this rank, completion
        event); candidate-overflow flags are checked in `drain()`."""
        from .extractor import extract_batch_device
        dev = torch.device('cuda', torch.cuda.current_device())
        main = torch.cuda.current_stream()
        if not hasattr(self, '_s_in'):
            self._s_in, self._s_out = torch.cuda.Stream(), torch.cuda.Stream()
        B, H, W = host_images.shape
        cap = host_out['x'].shape[1]
        key = (B, H, W, cap, dev.index)
        if getattr(self, '_slots_key', None) != key:
            ok = self.drain()                                   # the batches still in flight use the old slots
            self._overflow = not ok                             # ... and their overflow report must survive the re-key
            i32 = dict(dtype=torch.int32, device=dev)
            self._slots = [{'imgs': torch.empty((B, H, W), dtype=torch.float32, device=dev),
                            'x': torch.empty((B, cap), **i32), 'y': torch.empty((B, cap), **i32),
                            'count': torch.empty((B,), **i32),
                            'desc': torch.empty((B, cap, 128), dtype=torch.float32, device=dev),
                            'hflags': torch.zeros((B,), dtype=torch.int32).pin_memory(),
                            'imgs_free': None, 'out_free': None} for _ in range(2)]
            self._slots_key, self._slot_i, self._pending = key, 0, []
            main.synchronize()
        slot = self._slots[self._slot_i]
        self._slot_i ^= 1
        imgs = slot['imgs']
        full = {k: slot[k] for k in ('x', 'y', 'count', 'desc')}
        if slot['imgs_free'] is not None:
            self._s_in.wait_event(slot['imgs_free'])            # the extraction that last read this image slot is done
        if slot['out_free'] is not None:
            main.wait_event(slot['out_free'])                   # the read-back of this slot's previous results is done
        bounds = [(c0, min(c0 + chunk, B)) for c0 in range(0, B, chunk)]
        ready = []
        with torch.cuda.stream(self._s_in):
            for c0, c1 in bounds:
                imgs[c0:c1].copy_(host_images[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._s_in)
                ready.append(ev)
        hflags = slot['hflags']                                 # overflow flags land in pinned memory with the results:
        for ci, ((c0, c1), ev) in enumerate(zip(bounds, ready)):    # reading them later needs no device-wide wait
            main.wait_event(ev)
            res = extract_batch_device(imgs[c0:c1], self.params, want_aux=False, check=False,
                                       out={k: v[c0:c1] for k, v in full.items()})
            hflags[ci:ci + 1].copy_(res['_flag'], non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            self._s_out.wait_event(done)
            with torch.cuda.stream(self._s_out):
                for k in ('x', 'y', 'desc', 'count'):
                    host_out[k][c0:c1].copy_(full[k][c0:c1], non_blocking=True)
        slot['imgs_free'] = torch.cuda.Event()
        slot['imgs_free'].record(main)
        plan = self.pair_plan(pairs_global, B)                  # cached: the pair list rarely changes
        mine = plan.mine
        m = self.match_plan(plan, full['desc'], full['count'], cap=host_out['matches'].shape[1])
        if m is not None:
            host_out['matches'][:len(mine)].copy_(m[0], non_blocking=True)
            host_out['conf'][:len(mine)].copy_(m[1], non_blocking=True)
            host_out['mcount'][:len(mine)].copy_(m[2], non_blocking=True)
        main.wait_stream(self._s_out)
        slot['out_free'] = torch.cuda.Event()
        slot['out_free'].record(main)
        self._pending.append((slot['out_free'], hflags, len(bounds)))
        while len(self._pending) > 1:                           # a slot's flags are overwritten two calls later: retire
            ev, hf, nch = self._pending.pop(0)                  # the previous batch now (waits for that batch only)
            ev.synchronize()
            self._overflow = getattr(self, '_overflow', False) or bool(hf[:nch].any())
        return mine, slot['out_free']

    def drain(self) -> bool:
        """Wait for every batch given to `stream_host`; True when no candidate buffer overflowed (otherwise set
        `params.cand_full = 1` and resubmit the affected batches)."""
        ok = not getattr(self, '_overflow', False)
        for ev, hf, nch in getattr(self, '_pending', []):
            ev.synchronize()
            ok = ok and not bool(hf[:nch].any())
        self._pending, self._overflow = [], False
        return ok

    def pair_inliers(self, x_all: torch.Tensor, y_all: torch.Tensor, match, pairs_mine: np.ndarray, iterations: int,
                     threshold: float = 1.0, num_matches: int = 2500, threads: int = 8):
        """The reference's per-pair tail (Runner.py:347-351) for this rank's pairs, device-resident:
        matches -> coordinates (`sfm_matches_to_coords`) -> `find_inliers` (`sfm_find_inliers`).  Pairs are
        independent, so the stage shards with the pairs and needs no collective beyond the keypoint
        gather.  `match` is the tuple `self.match` returned for `pairs_mine`.  One host read of the match
        counts sizes the 8-subset draws, which run on `threads` host threads while the kernels queue.
        Returns a list, one entry per pair: None (fewer than 8 matches: the reference returns Nones), or
        (p1 [n,2] f64, p2, inlier_idx [n] i32, result [4] i32) device tensors."""
        from concurrent.futures import ThreadPoolExecutor
        from . import geometry as GE
        matches, _, mcount = match[0], match[1], match[2]
        counts = np.minimum(mcount.cpu().numpy(), num_matches)
        todo = [k for k in range(len(pairs_mine)) if counts[k] >= 8]
        out: List[Optional[tuple]] = [None] * len(pairs_mine)
        if not todo or iterations < 1:
            return out
        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
            samples = list(ex.map(lambda k: GE._samples_host(int(counts[k]), int(iterations), GE.RANSAC_SEED), todo))
        for k, smp in zip(todo, samples):
            i, j = int(pairs_mine[k][0]), int(pairs_mine[k][1])
            p1, p2, _ = GE.matches_to_coords_device(matches[k], mcount[k:k + 1], x_all[i], y_all[i], x_all[j], y_all[j],
                                                    num_matches)
            n = int(counts[k])
            idx, res, _ = GE.ransac_device(p1[:n], p2[:n], iterations, threshold,
                                           samples=smp.to(p1.device, non_blocking=True))
            out[k] = (p1[:n], p2[:n], idx, res)
        return out

    def step(self, images: torch.Tensor, pairs_global: np.ndarray):
        """Extract the local images, exchange the descriptor blocks the pair plan needs, match this
        rank's share of `pairs_global` (see PairPlan).  Returns (extraction outputs, matches, plan)."""
        out = self.extract(images)
        plan = self.pair_plan(pairs_global, images.shape[0])
        return out, self.match_plan(plan, out['desc'], out['count']), plan
