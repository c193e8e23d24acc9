"""uint8 min/max quantisation of frames (core/compressor.py:256-303) on the device.

`FrameQuantizer` keeps the reference's stateful contract (`_normalize_for_compression`
stores min/max on the instance, `_denormalize_from_compression` uses the last stored
pair); the batched functions carry min/max explicitly per frame."""
from __future__ import annotations

from typing import Optional, Tuple

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


def map_index_quantize(embeddings: torch.Tensor, n: Optional[int] = None, *, variant: str = "B", index_space: Optional[int] = None,
                       want_indices: bool = False):
    """map_to_2d + hierarchical index + embed + uint8 normalise of a batch in ONE kernel (hq_map_index_quant): what
    QuantizationPipeline.quantize_model hands to the image codec (core/pipeline.py:129-146, core/compressor.py:256-280).

    embeddings float32 [N, D] on the device.  Returns (frames uint8 [N, n + rows, n], minmax float32 [N, 2], indices or None):
    rows = 1 (variants A / B, index_space = n values, the pipeline's layout) or L (variant C).  Shapes the fused kernel does
    not cover (grids other than 32 x 32 / 64 x 64, unaligned rows) run the two-launch path (fused map + index into an
    enhanced float32 frame, then hq_quantize_u8) with identical results."""
    from . import plans
    from .index import _plan_tensor, fused_pass
    d = dev.require_cuda(embeddings.device)
    if embeddings.dtype != torch.float32 or embeddings.dim() != 2:
        raise TypeError("map_index_quantize expects float32 [N, D]")
    emb = embeddings.contiguous()
    N, D = emb.shape
    if n is None:
        from .dimension import rag_optimal_dimensions
        n = rag_optimal_dimensions(D)[0]
    pyr_mode = 0
    if variant == "C":
        plan, widths, ml = plans.c_plan(n, "rows")
        key = ("C", n, "rows")
    else:
        S = int(index_space if index_space is not None else n)
        if S != n:
            raise ValueError("the embedded index row holds exactly n values (core/pipeline.py:112, core/index_generator.py:241-245)")
        plan, ml = plans.a_plan(n, S) if variant == "A" else plans.b_plan(n, S)
        key, pyr_mode = ((variant, n, S), 0 if variant == "A" else 1)
    rows = len(plan) // n
    frames = torch.empty((N, n + rows, n), dtype=torch.uint8, device=d)
    mm = torch.empty((N, 2), dtype=torch.float32, device=d)
    idx = None
    if want_indices:
        idx = torch.empty((N, len(plan)), dtype=torch.float64 if pyr_mode else torch.float32, device=d)
    if N == 0:
        return frames, mm, idx
    plan_t = _plan_tensor(key, plan, d)
    frame_zero = int(bool((plan < 0).any()))
    from ._lib import HQ_EUNSUPPORTED
    with torch.cuda.device(d):
        src_stride = emb.stride(0) if N > 1 else D        # (a size-1 axis may report any stride)
        rc = lib.hq_map_index_quant(dev.ptr(emb), N, D, src_stride, n, dev.ptr(plan_t), len(plan), pyr_mode, ml, frame_zero,
                                    dev.ptr(frames), (n + rows) * n, dev.ptr(mm), dev.ptr(idx), len(plan), dev.stream_ptr())
    if rc == HQ_EUNSUPPORTED:
        # two launches: enhanced float32 frame (grid + index rows in the image dtype), then the frame quantiser
        enhanced = torch.empty((N, n + rows, n), dtype=torch.float32, device=d)
        flat = enhanced.view(N, -1)
        if pyr_mode:
            _, _, i64 = fused_pass(emb, 0, n, D, plan=plan, plan_key=key, min_level=ml, pyr_mode=1, grid_out=flat,
                                   grid_stride=(n + rows) * n)
            flat[:, n * n:] = i64.to(torch.float32)
            if idx is not None:
                idx.copy_(i64)
        else:
            fused_pass(emb, 0, n, D, plan=plan, plan_key=key, min_level=ml, grid_out=flat, grid_stride=(n + rows) * n,
                       idx_out=flat[:, n * n:], idx_stride=(n + rows) * n)
            if idx is not None:
                idx.copy_(flat[:, n * n:])
        q, mm = quantize_u8_batch(enhanced)
        return q, mm, idx
    check(rc)
    return frames, mm, idx


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
