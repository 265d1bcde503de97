"""Host-side mirror of the reference's FeatureExtractor classes.

Same class names, constructor arguments, return types and error behaviour as
FeatureExtractor/FeatureExtractor.py:4-21, FeatureExtractor/SIFT/NaiveSIFT.py:9-52
and FeatureExtractor/SIFT/ScaleRotInvSIFT.py:8-22 of reesque/SfmFromScratch, with
all arithmetic done by libsfmb200.so on a B200.  PyTorch is used only for
device buffers, pinned staging and the stream handle.
"""
from __future__ import annotations

import ctypes as C
import threading
from abc import ABC, abstractmethod
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _native as N


def _generate_gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """NaiveSIFT.py:175-199, evaluated with numpy exactly as the reference does
    so the float32 window weights handed to the kernel are the reference's."""
    mean = ksize // 2
    axis = np.linspace(-mean, mean, ksize)
    x_square = axis[:, np.newaxis] ** 2
    y_square = axis[np.newaxis, :] ** 2
    kernel = (1 / (2 * np.pi * sigma ** 2)) * np.exp(-(x_square + y_square) / (2 * sigma ** 2))
    return kernel / np.sum(kernel)


def make_params(extractor_params: Optional[dict], *, pyramid: bool):
    """extractor_params dict (main.py:19-28) -> (SfmExtractParams, keep-alive weights)."""
    ep = extractor_params or {}
    p = N.SfmExtractParams()
    N.load_library().sfm_extract_default_params(C.byref(p))
    p.num_interest_points = int(ep.get('num_interest_points', 2500))
    p.ksize = int(ep.get('ksize', 7))
    p.gaussian_size = int(ep.get('gaussian_size', 7))
    p.sigma = float(ep.get('sigma', 5))
    p.alpha = float(ep.get('alpha', 0.05))
    p.feature_width = int(ep.get('feature_width', 16))
    if pyramid:
        p.pyramid_level = int(ep.get('pyramid_level', 4))
        p.pyramid_scale_factor = float(ep.get('pyramid_scale_factor', 2))
        p.rotation_invariant = 1
        p.split_k_by_level = 1
    else:
        p.pyramid_level = 1
        p.pyramid_scale_factor = 1.0
        p.rotation_invariant = 0
        p.split_k_by_level = 0
    w = np.ascontiguousarray(_generate_gaussian_kernel(p.gaussian_size, p.sigma), dtype=np.float32)
    p.gauss_weights = w.ctypes.data_as(C.POINTER(C.c_float))
    return p, w


def copy_params(p: N.SfmExtractParams) -> N.SfmExtractParams:
    """A private copy of a parameter struct (the weight pointer is shared; keep its owner alive)."""
    return N.SfmExtractParams.from_buffer_copy(p)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def extract_batch_device(images: torch.Tensor, params: N.SfmExtractParams, *, want_aux: bool = True,
                         device_index: Optional[int] = None, out: Optional[Dict[str, torch.Tensor]] = None,
                         check: bool = True) -> Dict[str, torch.Tensor]:
    """Run sfm_extract_batch on a float32 CUDA tensor [B, H, W].  Returns device
    tensors: x, y, count, desc (+ lx, ly, level, conf when want_aux).  `out` may
    hold preallocated (contiguous) result tensors, e.g. batch slices of larger
    ones.  Asynchronous on the current stream; with check=True (default) the
    candidate-overflow flag is read back (a 4-byte copy and a stream sync) and a
    plateau image is retried with full-size buffers, with check=False the caller
    does that later through check_extract_status(result)."""
    if not images.is_cuda or images.dtype != torch.float32 or images.dim() != 3:
        raise ValueError("images must be a float32 CUDA tensor of shape [B, H, W]")
    images = images.contiguous()
    dev = images.device
    idx = dev.index if device_index is None else device_index
    L = N.load_library()
    ctx = N.get_ctx(idx)
    B, H, W = images.shape
    with torch.cuda.device(dev):
        cap = L.sfm_extract_max_keypoints(C.byref(params))
        i32 = dict(dtype=torch.int32, device=dev)
        if out is None:
            out = {
                'x': torch.empty((B, cap), **i32), 'y': torch.empty((B, cap), **i32),
                'count': torch.empty((B,), **i32),
                'desc': torch.empty((B, cap, N.DESC_DIM), dtype=torch.float32, device=dev),
            }
            if want_aux:
                out.update(lx=torch.empty((B, cap), **i32), ly=torch.empty((B, cap), **i32),
                           level=torch.empty((B, cap), **i32),
                           conf=torch.empty((B, cap), dtype=torch.float32, device=dev))
        else:
            for k in ('x', 'y', 'count', 'desc'):
                if not out[k].is_contiguous() or out[k].shape[0] != B:
                    raise ValueError(f"out[{k!r}] must be contiguous with leading dimension {B}")
            cap = out['x'].shape[1]
        ptr = lambda k: out[k].data_ptr() if k in out else None
        # a plateau image is retried on a private copy of the parameters: the caller's struct may be shared
        # between threads (Runner.py:186-191 runs extractors from an 8-thread pool) and must not change under them
        params = copy_params(params)
        for attempt in range(2):
            nbytes = L.sfm_extract_workspace_bytes(B, H, W, C.byref(params))
            if nbytes == 0:
                # let the library produce the precise message
                N.check(L.sfm_extract_batch(ctx, _stream_ptr(), images.data_ptr(), B, H, W, C.byref(params),
                                            None, 0, ptr('x'), ptr('y'), None, None, None, None,
                                            ptr('desc'), ptr('count'), cap), ctx)
                raise RuntimeError("sfm_extract_workspace_bytes returned 0")
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            N.check(L.sfm_extract_batch(ctx, _stream_ptr(), images.data_ptr(), B, H, W, C.byref(params),
                                        ws.data_ptr(), nbytes, ptr('x'), ptr('y'), ptr('lx'), ptr('ly'),
                                        ptr('level'), ptr('conf'), ptr('desc'), ptr('count'), cap), ctx)
            if not check:
                out['_flag'] = ws[:4].view(torch.int32).clone()    # overflow flag, read by check_extract_status
                break
            rc = L.sfm_extract_status(ctx, _stream_ptr(), ws.data_ptr())
            if rc == N.SFM_ERR_CAPACITY and attempt == 0 and not params.cand_full:
                params.cand_full = 1          # plateau image: one candidate slot per pixel
                continue
            N.check(rc, ctx)
            break
    return out


def check_extract_status(result: Dict[str, torch.Tensor]) -> bool:
    """For check=False calls: True when no candidate buffer overflowed (the
    result is valid); False means re-run with params.cand_full = 1."""
    f = result.get('_flag')
    return True if f is None else int(f.cpu()[0]) == 0


_staging = threading.local()


def _to_device(image: np.ndarray) -> torch.Tensor:
    """numpy image -> CUDA tensor through a pinned staging buffer kept per thread and shape (pinning a fresh
    buffer on every class call cost more than the extraction itself: cudaHostAlloc is a device-wide sync)."""
    a = np.ascontiguousarray(image, dtype=np.float32)
    cache = getattr(_staging, 'bufs', None)
    if cache is None:
        cache = _staging.bufs = {}
    ent = cache.get(a.shape)
    if ent is None:
        if len(cache) >= 4:                       # a handful of shapes per thread is the working set of a run
            cache.pop(next(iter(cache)))
        ent = cache[a.shape] = [torch.empty(a.shape, dtype=torch.float32).pin_memory(), None]
    buf, ev = ent
    if ev is not None:
        ev.synchronize()                          # the previous copy out of this buffer has finished
    buf.numpy()[...] = a
    dev = buf.to('cuda', non_blocking=True)
    ent[1] = torch.cuda.Event()
    ent[1].record()
    return dev


class FeatureExtractor(ABC):
    """FeatureExtractor/FeatureExtractor.py:4-21."""

    def __init__(self, image: np.ndarray, extractor_params=None):
        if extractor_params is None:
            extractor_params = {}
        self.image = image
        self.num_interest_points = extractor_params.get('num_interest_points', 2500)

    @abstractmethod
    def detect_keypoints(self) -> np.ndarray:
        pass

    @abstractmethod
    def extract_descriptors(self) -> np.ndarray:
        pass


class NaiveSIFT(FeatureExtractor):
    """FeatureExtractor/SIFT/NaiveSIFT.py:9-52: single-scale Harris keypoints and
    unrotated 4x4x8 descriptors, computed lazily by detect_keypoints()."""

    def __init__(self, image_bw: np.ndarray, extractor_params: dict = {}):
        self.SOBEL_X_KERNEL = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]).astype(np.float32)
        self.SOBEL_Y_KERNEL = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]]).astype(np.float32)
        super().__init__(image_bw, extractor_params)
        self._params_dict = dict(extractor_params or {})
        self._ksize = extractor_params.get('ksize', 7)
        self._gaussian_size = extractor_params.get('gaussian_size', 7)
        self._sigma = extractor_params.get('sigma', 5)
        self._alpha = extractor_params.get('alpha', 0.05)
        self._feature_width = extractor_params.get('feature_width', 16)

    def _run(self, pyramid: bool):
        assert np.ndim(self.image) == 2, 'Image must be grayscale'       # NaiveSIFT.py:127
        p, keep = make_params(self._params_dict, pyramid=pyramid)
        out = extract_batch_device(_to_device(self.image)[None], p)
        n = int(out['count'].cpu()[0])
        host = {k: v[0, :n].cpu().numpy() for k, v in out.items() if k != 'count'}
        del keep
        return n, host

    def detect_keypoints(self) -> Tuple[np.ndarray, np.ndarray]:
        n, h = self._run(pyramid=False)
        self._X = h['x'].astype(np.int64)
        self._Y = h['y'].astype(np.int64)
        self.confidences = h['conf']
        self._descriptors = h['desc']
        return self._X, self._Y

    def extract_descriptors(self) -> np.ndarray:
        if not hasattr(self, '_X') or not hasattr(self, '_Y'):
            raise RuntimeError("Keypoints not detected. Call detect_keypoints() before extract_descriptors().")
        # NaiveSIFT.py:173 returns np.squeeze(np.array(fvs)): (n,128), (128,) for n == 1, (0,) for n == 0
        d = self._descriptors
        self.descriptors = np.squeeze(d) if len(d) else np.array([])
        return self.descriptors

    def _generate_gaussian_kernel(self, ksize: int, sigma: float) -> np.ndarray:
        return _generate_gaussian_kernel(ksize, sigma)


class ScaleRotInvSIFT(NaiveSIFT):
    """FeatureExtractor/SIFT/ScaleRotInvSIFT.py:8-22: image pyramid, per-level
    Harris keypoints, rotation-normalised descriptors; all work happens in the
    constructor, the two methods are getters."""

    def __init__(self, image_bw: np.ndarray, extractor_params: dict = {}):
        super().__init__(image_bw, extractor_params)
        self._pyramid_level = extractor_params.get('pyramid_level', 4)
        self._pyramid_scale_factor = extractor_params.get('pyramid_scale_factor', 2)
        self.compute(self.num_interest_points)

    def detect_keypoints(self):
        return self._X, self._Y

    def extract_descriptors(self):
        return self._feature_vec

    def compute(self, k: int):
        """ScaleRotInvSIFT.py:89-107 on the GPU."""
        self._params_dict['num_interest_points'] = k
        n, h = self._run(pyramid=True)
        if n == 0:
            # np.array([]) of the reference's empty lists
            self._X, self._Y, self._feature_vec = np.array([]), np.array([]), np.array([])
        else:
            self._X = h['x'].astype(np.int64)
            self._Y = h['y'].astype(np.int64)
            self._feature_vec = h['desc']
        # extras for diagnostics / parity tests (not part of the reference API)
        self.levels = h['level'].astype(np.int64)
        self.level_x = h['lx'].astype(np.int64)
        self.level_y = h['ly'].astype(np.int64)
        self.confidences = h['conf']


def extract_batch(images, extractor_params: Optional[dict] = None, *, pyramid: bool = True):
    """Batched extraction.  `images`: numpy [B,H,W] (copied through pinned
    memory) or a float32 CUDA tensor.  Returns a list of (X, Y, descriptors)
    numpy triples shaped like the reference's per-image results."""
    p, keep = make_params(extractor_params, pyramid=pyramid)
    if isinstance(images, np.ndarray):
        dev = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).pin_memory().to('cuda', non_blocking=True)
    else:
        dev = images
    out = extract_batch_device(dev, p, want_aux=False)
    counts = out['count'].cpu().numpy()
    x, y, d = out['x'].cpu().numpy(), out['y'].cpu().numpy(), out['desc'].cpu().numpy()
    del keep
    return [(x[b, :n].astype(np.int64), y[b, :n].astype(np.int64), d[b, :n]) for b, n in enumerate(counts)]


def harris_response(image: np.ndarray, extractor_params: Optional[dict] = None) -> np.ndarray:
    """R map of NaiveSIFT.py:60-74 through sfm_harris_response (parity tests)."""
    p, keep = make_params(extractor_params, pyramid=False)
    img = _to_device(image)
    H, W = img.shape
    out = torch.empty((H, W), dtype=torch.float32, device=img.device)
    L = N.load_library()
    ctx = N.get_ctx(img.device.index)
    N.check(L.sfm_harris_response(ctx, _stream_ptr(), img.data_ptr(), H, W, C.byref(p), out.data_ptr()), ctx)
    res = out.cpu().numpy()
    del keep
    return res
