"""Device plumbing shared by the host classes: torch owns memory and streams, the C ABI
gets raw pointers.  No compute happens here."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import HQLibraryError

_WIDTH_TO_TORCH = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
_WIDTH_TO_NUMPY = {1: np.uint8, 2: np.int16, 4: np.int32, 8: np.int64}


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise HQLibraryError("hilbert_quantization_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise HQLibraryError(f"device {dev} is not a CUDA device; there is no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_words(arr: np.ndarray) -> np.ndarray:
    """View an array as same-width integers (pure byte moves preserve any dtype)."""
    w = arr.dtype.itemsize
    if w not in _WIDTH_TO_NUMPY:
        raise TypeError(f"unsupported element width {w} bytes ({arr.dtype})")
    return np.ascontiguousarray(arr).view(_WIDTH_TO_NUMPY[w])


def to_device(arr: np.ndarray, device: torch.device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=False)


def canonical_strides(t: torch.Tensor) -> torch.Tensor:
    """A contiguous tensor whose size-1 axes also report row-major strides (NumPy and torch keep whatever stride such an
    axis had, e.g. `q[None, :]`; the C ABI takes `stride(0)` at face value)."""
    if t.dim() >= 2 and t.numel() > 0 and t.is_contiguous():
        return t.view(t.numel()).view(t.shape)
    return t


def f32_device(x, device: torch.device) -> torch.Tensor:
    """float32, contiguous, on `device` (accepts ndarray or tensor)."""
    if isinstance(x, torch.Tensor):
        return canonical_strides(x.to(device=device, dtype=torch.float32).contiguous())
    return canonical_strides(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device))
