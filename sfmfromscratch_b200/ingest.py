"""Image ingest in front of the feature path: the host-side mirror
of Runner.py:33-46 -- `_load_image` (:551-563), `_PIL_resize` (:481-493), `_rgb2gray` (:467-478).

JPEG/PNG decoding stays with PIL on the host; the resize (PIL's default BICUBIC on RGB), the
float32/255 conversion and the grayscale mix run on the GPU (sfm_ingest_rgb8) and return the
float32 array the reference would hand to its extractor, bit for bit.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _native as N


def gray_from_rgb8_device(rgb: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """rgb: uint8 CUDA tensor [B, H, W, 3] -> float32 CUDA tensor [B, out_h, out_w]."""
    if not rgb.is_cuda or rgb.dtype != torch.uint8 or rgb.dim() != 4 or rgb.shape[3] != 3:
        raise ValueError("rgb must be a uint8 CUDA tensor of shape [B, H, W, 3]")
    rgb = rgb.contiguous()
    B, H, W, _ = rgb.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    L = N.load_library()
    ctx = N.get_ctx(rgb.device.index)
    with torch.cuda.device(rgb.device):
        nbytes = L.sfm_ingest_workspace_bytes(B, H, W, oh, ow)
        if nbytes == 0:
            raise ValueError(f"bad ingest sizes {(B, H, W)} -> {(oh, ow)}")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=rgb.device)
        out = torch.empty((B, oh, ow), dtype=torch.float32, device=rgb.device)
        N.check(L.sfm_ingest_rgb8(ctx, torch.cuda.current_stream().cuda_stream, rgb.data_ptr(), B, H, W, oh, ow,
                                  ws.data_ptr(), nbytes, out.data_ptr()), ctx)
        ws.record_stream(torch.cuda.current_stream())
    return out


def gray_from_rgb8(img_u8: np.ndarray, scale_factor: float = 0.5) -> np.ndarray:
    """Decoded 8-bit RGB image (H, W, 3) -> the reference's `_image_bw`: float32 (int(H*s), int(W*s))."""
    img_u8 = np.ascontiguousarray(img_u8)
    if img_u8.dtype != np.uint8 or img_u8.ndim != 3 or img_u8.shape[2] != 3:
        raise ValueError("expected an (H, W, 3) uint8 RGB image (the reference's _rgb2gray indexes three channels)")
    H, W, _ = img_u8.shape
    size = (int(W * scale_factor), int(H * scale_factor))          # Runner.py:38-40: (width, height)
    dev = torch.from_numpy(img_u8).pin_memory().to('cuda', non_blocking=True)[None]
    return gray_from_rgb8_device(dev, (size[1], size[0]))[0].cpu().numpy()


def load_image_gray(path: str, scale_factor: float = 0.5) -> np.ndarray:
    """Runner.py:33-46 for one image file."""
    import PIL.Image
    img = PIL.Image.open(path)
    arr = np.asarray(img)
    if arr.dtype != np.uint8:
        raise ValueError("only 8-bit images are supported")
    return gray_from_rgb8(arr, scale_factor)
