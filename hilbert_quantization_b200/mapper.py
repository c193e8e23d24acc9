"""HilbertCurveMapper / HilbertCurveMapperImpl backed by the sm_100a kernels.

Interface mirrored: hilbert_quantization/interfaces.py:43-84, implementation semantics
core/hilbert_mapper.py:14-205 and rag/embedding_generation/hilbert_mapper.py:9-230
(same names, argument order, return types, exception types and message texts).
Single items are NumPy in / NumPy out (H2D -> kernel -> D2H); the batched `*_batch`
methods take and return device tensors and are what the benchmarks time.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib
from .exceptions import HilbertQuantizationError


def _valid_side(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


class _MapperCore:
    """Shared device calls; subclasses choose the exception type."""
    _exc = HilbertQuantizationError

    def __init__(self, device=None):
        self._device = device          # resolved lazily so construction works without a GPU

    @property
    def device(self) -> torch.device:
        d = dev.require_cuda(self._device)
        self._device = d
        return d

    # ---- coordinates -------------------------------------------------
    def _coords_device(self, n: int):
        d = self.device
        with torch.cuda.device(d):
            x = torch.empty(n * n, dtype=torch.int32, device=d)
            y = torch.empty(n * n, dtype=torch.int32, device=d)
            check(lib.hq_d2xy_batch(n, 0, n * n, dev.ptr(x), dev.ptr(y), dev.stream_ptr()), self._exc)
        return x, y

    def generate_hilbert_coordinates(self, n: int) -> List[Tuple[int, int]]:
        if n <= 0 or (n & (n - 1)) != 0:
            raise self._exc(f"Grid size must be a power of 2, got {n}")
        x, y = self._coords_device(n)
        return list(zip(x.cpu().tolist(), y.cpu().tolist()))

    def _hilbert_index_to_xy(self, index: int, n: int) -> Tuple[int, int]:
        d = self.device
        with torch.cuda.device(d):
            x = torch.empty(1, dtype=torch.int32, device=d)
            y = torch.empty(1, dtype=torch.int32, device=d)
            check(lib.hq_d2xy_batch(n, int(index), 1, dev.ptr(x), dev.ptr(y), dev.stream_ptr()), self._exc)
        return int(x.item()), int(y.item())

    def _xy_to_hilbert_index(self, x: int, y: int, n: int) -> int:
        d = self.device
        with torch.cuda.device(d):
            xs = torch.tensor([x], dtype=torch.int32, device=d)
            ys = torch.tensor([y], dtype=torch.int32, device=d)
            out = torch.empty(1, dtype=torch.int64, device=d)
            check(lib.hq_xy2d_batch(n, dev.ptr(xs), dev.ptr(ys), 1, dev.ptr(out), dev.stream_ptr()), self._exc)
        return int(out.item())

    def _rotate(self, n: int, x: int, y: int, rx: int, ry: int) -> Tuple[int, int]:
        """core/hilbert_mapper.py:92-113 (scalar helper kept for API parity)."""
        if ry == 0:
            if rx == 1:
                x = n - 1 - x
                y = n - 1 - y
            x, y = y, x
        return x, y

    # ---- batched device path -----------------------------------------
    def map_to_2d_batch(self, params: torch.Tensor, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[N, D] device tensor (any 1/2/4/8-byte dtype) -> [N, n, n], zero padded."""
        if params.dim() != 2:
            raise self._exc(f"expected [N, D], got shape {tuple(params.shape)}")
        if not _valid_side(n):
            raise self._exc(f"Dimension must be a power of 2, got {n}")
        N, D = params.shape
        if D > n * n:
            raise self._exc(f"Too many parameters ({D}) for dimensions {n}x{n} ({n * n} cells)")
        d = dev.require_cuda(params.device)
        params = params.contiguous()
        if out is None:
            out = torch.empty((N, n, n), dtype=params.dtype, device=d)
        with torch.cuda.device(d):
            check(lib.hq_map_to_2d(dev.ptr(params), N, D, params.stride(0) if N else D, n, params.element_size(),
                                   dev.ptr(out), n * n, dev.stream_ptr()), self._exc)
        return out

    def map_from_2d_batch(self, grids: torch.Tensor, length: Optional[int] = None,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[N, n, n] device tensor -> [N, length] (default n*n) in curve order."""
        if grids.dim() != 3 or grids.shape[1] != grids.shape[2]:
            raise self._exc(f"Hilbert curve requires square dimensions, got {tuple(grids.shape)}")
        N, n, _ = grids.shape
        if not _valid_side(n):
            raise self._exc(f"Dimension must be a power of 2, got {n}")
        L = n * n if length is None else int(length)
        d = dev.require_cuda(grids.device)
        grids = grids.contiguous()
        if out is None:
            out = torch.empty((N, L), dtype=grids.dtype, device=d)
        with torch.cuda.device(d):
            check(lib.hq_map_from_2d(dev.ptr(grids), N, n, n * n, L, grids.element_size(), dev.ptr(out), L,
                                     dev.stream_ptr()), self._exc)
        return out

    # ---- single item, NumPy in / NumPy out ----------------------------
    def _map_to_2d_numpy(self, parameters: np.ndarray, width: int) -> np.ndarray:
        parameters = np.asarray(parameters)
        words = dev.as_words(parameters.reshape(-1))
        t = dev.to_device(words, self.device).reshape(1, -1)
        if t.shape[1] == 0:
            return np.zeros((width, width), dtype=parameters.dtype)
        g = self.map_to_2d_batch(t, width)
        return g[0].cpu().numpy().view(parameters.dtype)

    def _map_from_2d_numpy(self, image: np.ndarray) -> np.ndarray:
        image = np.asarray(image)
        n = image.shape[0]
        words = dev.as_words(image)
        t = dev.to_device(words, self.device).reshape(1, n, n)
        s = self.map_from_2d_batch(t)
        return s[0].cpu().numpy().view(image.dtype)


class HilbertCurveMapper(_MapperCore):
    """Drop-in for hilbert_quantization.core.hilbert_mapper.HilbertCurveMapper."""

    def map_to_2d(self, parameters: np.ndarray, dimensions: Tuple[int, int], builder=None) -> np.ndarray:
        width, height = dimensions
        if width != height:
            raise HilbertQuantizationError(f"Hilbert curve requires square dimensions, got {width}x{height}")
        if width <= 0 or (width & (width - 1)) != 0:
            raise HilbertQuantizationError(f"Dimension must be a power of 2, got {width}")
        total_cells = width * height
        if len(parameters) > total_cells:
            raise HilbertQuantizationError(
                f"Too many parameters ({len(parameters)}) for dimensions {width}x{height} ({total_cells} cells)")
        result = self._map_to_2d_numpy(parameters, width)
        if builder is not None:
            # core/hilbert_mapper.py:152-172: the optional builder is fed the real values in order
            if hasattr(builder, "reset"):
                builder.reset()
            feed = getattr(builder, "add_parameter_value", None) or getattr(builder, "add_value", None)
            if feed is not None:
                for v in np.asarray(parameters).reshape(-1)[:total_cells]:
                    feed(float(v))
        return result

    def map_from_2d(self, image: np.ndarray) -> np.ndarray:
        height, width = image.shape
        if width != height:
            raise HilbertQuantizationError(f"Hilbert curve requires square dimensions, got {width}x{height}")
        if width <= 0 or (width & (width - 1)) != 0:
            raise HilbertQuantizationError(f"Dimension must be a power of 2, got {width}")
        return self._map_from_2d_numpy(image)


class HilbertCurveMapperImpl(_MapperCore):
    """Drop-in for rag.embedding_generation.hilbert_mapper.HilbertCurveMapperImpl (ValueError flavour)."""
    _exc = ValueError

    def __init__(self, config=None, device=None):
        super().__init__(device)
        self.config = config

    def map_to_2d(self, embeddings: np.ndarray, dimensions: Tuple[int, int]) -> np.ndarray:
        width, height = dimensions
        if width <= 0 or height <= 0:
            raise ValueError(f"Dimensions must be positive, got {width}x{height}")
        if width != height:
            raise ValueError(f"Hilbert curve requires square dimensions, got {width}x{height}")
        if width <= 0 or (width & (width - 1)) != 0:
            raise ValueError(f"Dimension must be a power of 2, got {width}")
        total_cells = width * height
        if len(embeddings) > total_cells:
            raise ValueError(
                f"Too many embedding values ({len(embeddings)}) for dimensions {width}x{height} ({total_cells} cells)")
        return self._map_to_2d_numpy(embeddings, width)

    def map_from_2d(self, image: np.ndarray) -> np.ndarray:
        if len(image.shape) != 2:
            raise ValueError(f"Input must be 2D array, got {len(image.shape)}D")
        height, width = image.shape
        if width != height:
            raise ValueError(f"Hilbert curve requires square dimensions, got {width}x{height}")
        if width <= 0 or (width & (width - 1)) != 0:
            raise ValueError(f"Dimension must be a power of 2, got {width}")
        return self._map_from_2d_numpy(image)
