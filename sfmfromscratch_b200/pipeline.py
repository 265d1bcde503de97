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


class FeaturePipeline:
    """Extraction of this rank's image shard, descriptor exchange, matching of
    this rank's pair share.  All tensors stay on the device."""

    def __init__(self, extractor_params: Optional[dict] = None, ratio_threshold: float = 0.8,
                 rank: int = 0, world: int = 1, group=None):
        from .extractor import make_params
        self.params, self._keep = make_params(extractor_params, pyramid=True)
        self.ratio_threshold = ratio_threshold
        self.rank, self.world, self.group = rank, world, group

    def extract(self, images: torch.Tensor):
        from .extractor import extract_batch_device
        return extract_batch_device(images, self.params, want_aux=False)

    def exchange(self, desc: torch.Tensor, counts: torch.Tensor):
        return gather_descriptors(desc, counts, self.group)

    def match(self, desc_all: torch.Tensor, counts_all: torch.Tensor, pairs: torch.Tensor, cap: Optional[int] = None):
        from .matcher import match_batch_device
        if pairs.shape[0] == 0:
            return None
        return match_batch_device(desc_all, counts_all, pairs, self.ratio_threshold, cap=cap)

    def step(self, images: torch.Tensor, pairs_global: np.ndarray):
        """Extract the local images, all-gather, match this rank's share of `pairs_global`."""
        out = self.extract(images)
        desc_all, counts_all = self.exchange(out['desc'], out['count'])
        mine = deal_pairs(pairs_global, self.rank, self.world)
        pairs = torch.from_numpy(np.ascontiguousarray(mine)).to(images.device)
        return out, self.match(desc_all, counts_all, pairs)
