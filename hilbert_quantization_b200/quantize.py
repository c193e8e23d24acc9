"""uint8 min/max quantisation of frames (core/compressor.py:256-303) on the device.

`FrameQuantizer` keeps the reference's stateful contract (`_normalize_for_compression`
stores min/max on the instance, `_denormalize_from_compression` uses the last stored
pair); the batched functions carry min/max explicitly per frame."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib


def quantize_u8_batch(frames: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """float32 [N, ...] -> (uint8 same shape, minmax float32 [N, 2]); truncating cast, constant frame -> 128."""
    d = dev.require_cuda(frames.device)
    if frames.dtype != torch.float32:
        raise TypeError("quantize_u8_batch expects float32 frames")
    frames = frames.contiguous()
    N = frames.shape[0]
    elems = frames[0].numel() if N else 1
    out = torch.empty(frames.shape, dtype=torch.uint8, device=d)
    mm = torch.empty((N, 2), dtype=torch.float32, device=d)
    with torch.cuda.device(d):
        check(lib.hq_quantize_u8(dev.ptr(frames), N, elems, elems, dev.ptr(out), elems, dev.ptr(mm), dev.stream_ptr()))
    return out, mm


def dequantize_u8_batch(q: torch.Tensor, minmax: torch.Tensor) -> torch.Tensor:
    d = dev.require_cuda(q.device)
    q = q.contiguous()
    N = q.shape[0]
    elems = q[0].numel() if N else 1
    out = torch.empty(q.shape, dtype=torch.float32, device=d)
    mm = minmax.to(device=d, dtype=torch.float32).contiguous()
    with torch.cuda.device(d):
        check(lib.hq_dequantize_u8(dev.ptr(q), N, elems, elems, dev.ptr(mm), dev.ptr(out), elems, dev.stream_ptr()))
    return out


class FrameQuantizer:
    """The normalise / denormalise half of MPEGAICompressorImpl (the JPEG codec stays host plumbing)."""

    def __init__(self, device=None):
        self._device = device

    def _normalize_for_compression(self, image: np.ndarray) -> np.ndarray:
        d = dev.require_cuda(self._device)
        t = dev.f32_device(np.asarray(image), d).reshape(1, *image.shape)
        q, mm = quantize_u8_batch(t)
        mn, mx = (np.float32(v) for v in mm[0].cpu().tolist())
        if mx != mn:                       # compressor.py:268-278: constant images do not update the stored range
            self._norm_min, self._norm_max = mn, mx
        return q[0].cpu().numpy()

    def _denormalize_from_compression(self, image: np.ndarray) -> np.ndarray:
        if not hasattr(self, "_norm_min") or not hasattr(self, "_norm_max"):
            return image.astype(np.float32) / 255.0
        d = dev.require_cuda(self._device)
        t = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(d).reshape(1, *image.shape)
        mm = torch.tensor([[float(self._norm_min), float(self._norm_max)]], dtype=torch.float32)
        return dequantize_u8_batch(t, mm)[0].cpu().numpy()
