"""PrecomputedHilbertIndexer on the device (core/precomputed_hilbert_index.py:52-262).

The reference averages every aligned s x s square (s = 2, 4, ...) and every square offset by s / 2 with
Python loops over NumPy slices -- on every HilbertQuantizer.quantize call (api.py:162-173).  Aligned
power-of-two squares are contiguous runs of the Hilbert stream, so all of them come out of ONE fused
pyramid pass (hq_map_index_fused, direction 1) through a gather plan in the reference's row-major
order; the half-offset squares are four entries of the next finer level each (hq_offset_square_means).
"""
from __future__ import annotations

import math
import pickle
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _device as dev
from . import plans
from ._lib import check, lib
from .index import fused_pass
from .mapper import HilbertCurveMapper


@dataclass
class PrecomputedLevel:
    """core/precomputed_hilbert_index.py:23-30"""
    grid_size: int
    square_size: int
    num_squares: int
    averages: np.ndarray
    square_coordinates: List[Tuple[int, int]]


@dataclass
class PrecomputedIndex:
    """core/precomputed_hilbert_index.py:33-49"""
    model_id: str
    original_shape: Tuple[int, int]
    levels: List[PrecomputedLevel]
    creation_time: float
    total_storage_bytes: int


def _aligned_plan(n: int, square: int) -> np.ndarray:
    """Gather plan of the aligned square means of side `square` in row-major (row, col) order."""
    g = n // square
    rows, cols = np.divmod(np.arange(g * g), g)
    if square == 1:
        return (rows * n + cols).astype(np.int32)
    k = int(math.log2(square))
    return (plans.level_base(n, k) + plans._xy2d(g, cols, rows)).astype(np.int32)


class PrecomputedHilbertIndexer:
    """Same constructor, methods and result types as the reference class; the averages are float32 like
    the reference's (`np.mean` of a float32 region), computed in a 4-ary tree instead of NumPy's pairwise
    order (|delta| <= 3e-7 on N(0,1) data, tests)."""

    _level_cls = _index_cls = None   # dropin.install() points these at the reference's dataclasses (pickle compatibility)

    def __init__(self, max_levels: int = 6, min_square_size: int = 2, device=None):
        self.max_levels = max_levels
        self.min_square_size = min_square_size
        self.hilbert_mapper = HilbertCurveMapper()
        self._index_cache: Dict[str, PrecomputedIndex] = {}
        self._device = device

    def _calculate_granularity_levels(self, image_size: int) -> List[Tuple[int, int]]:
        """core/precomputed_hilbert_index.py:122-150"""
        levels = []
        square_size = self.min_square_size
        while square_size <= image_size // 2 and len(levels) < self.max_levels:
            grid_size = image_size // square_size
            if grid_size >= 2:
                levels.append((grid_size, square_size))
            square_size *= 2
        if len(levels) == 0 or levels[-1][1] < image_size:
            levels.append((1, image_size))
        return levels

    # ------------------------------------------------------------------ batch path (device tensors in / out)
    def level_averages_batch(self, grids: torch.Tensor) -> List[torch.Tensor]:
        """grids float32 [N, n, n] on the device -> one [N, g*g + (g-1)*(g-1)] tensor per level (aligned squares
        row-major, then the half-offset squares row-major), the order of `_precompute_level_averages`."""
        d = dev.require_cuda(grids.device)
        N, n, n2 = grids.shape
        if n != n2 or not plans.is_pow2(n) or n < 4:
            raise ValueError(f"Image must be a power-of-two square of side >= 4, got {n}x{n2}")
        levels = self._calculate_granularity_levels(n)
        for _, s in levels:
            if not plans.is_pow2(s):
                raise NotImplementedError("min_square_size must be a power of two")
        # every square size that is needed either as a level or as the finer level of an offset pass
        sizes = sorted({s for _, s in levels} | {s // 2 for g, s in levels if g >= 2})
        plan = np.concatenate([_aligned_plan(n, s) for s in sizes])
        offs = np.cumsum([0] + [(n // s) ** 2 for s in sizes])
        ml = min(int(math.log2(s)) for s in sizes if s > 1) if any(s > 1 for s in sizes) else 99
        flat = grids.contiguous().view(N, n * n)
        _, _, vals = fused_pass(flat, 1, n, n * n, plan=plan, plan_key=("P", n, tuple(sizes)), min_level=ml)
        out = []
        with torch.cuda.device(d):
            for g, s in levels:
                i = sizes.index(s)
                aligned = vals[:, offs[i]:offs[i + 1]]
                if g < 2:
                    out.append(aligned.contiguous())
                    continue
                j = sizes.index(s // 2)
                G = n // (s // 2)
                res = torch.empty((N, g * g + (g - 1) * (g - 1)), dtype=torch.float32, device=d)
                res[:, : g * g] = aligned
                if g > 1 and (g - 1) > 0:
                    half = vals[:, offs[j]:offs[j + 1]]
                    tail = res[:, g * g:]
                    check(lib.hq_offset_square_means(dev.ptr(half), N, G, vals.stride(0), dev.ptr(tail), res.stride(0), dev.stream_ptr()))
                out.append(res)
        return out

    @staticmethod
    def _coordinates(g: int, s: int) -> List[Tuple[int, int]]:
        coords = [(c * s, r * s) for r in range(g) for c in range(g)]
        off = s // 2
        if off > 0:
            coords += [(c * s + off, r * s + off) for r in range(g - 1) for c in range(g - 1)]
        return coords

    # ------------------------------------------------------------------ reference surface
    def _precompute_level_averages(self, image: np.ndarray, grid_size: int, square_size: int) -> PrecomputedLevel:
        idx = self.create_precomputed_index(image, "__level__", cache=False)
        for lvl in idx.levels:
            if lvl.grid_size == grid_size and lvl.square_size == square_size:
                return lvl
        raise ValueError(f"({grid_size}, {square_size}) is not a level of a {image.shape[0]}x{image.shape[1]} image")

    def create_precomputed_index(self, image: np.ndarray, model_id: str, cache: bool = True) -> PrecomputedIndex:
        """core/precomputed_hilbert_index.py:65-120"""
        start = time.time()
        height, width = image.shape
        if height != width:
            raise ValueError(f"Image must be square, got {height}x{width}")
        d = dev.require_cuda(self._device)
        grids = dev.f32_device(np.asarray(image, dtype=np.float32)[None], d)
        per_level = self.level_averages_batch(grids)
        levels, total = [], 0
        for (g, s), t in zip(self._calculate_granularity_levels(width), per_level):
            avg = t[0].cpu().numpy().astype(np.float32)
            coords = self._coordinates(g, s)
            levels.append((self._level_cls or PrecomputedLevel)(grid_size=g, square_size=s, num_squares=len(avg), averages=avg, square_coordinates=coords))
            total += avg.nbytes + len(coords) * 16
        index = (self._index_cls or PrecomputedIndex)(model_id=model_id, original_shape=(height, width), levels=levels,
                                 creation_time=time.time() - start, total_storage_bytes=total)
        if cache:
            self._index_cache[model_id] = index
        return index

    def get_index(self, model_id: str) -> Optional[PrecomputedIndex]:
        return self._index_cache.get(model_id)

    def save_index_to_disk(self, index: PrecomputedIndex, filepath: str):
        with open(filepath, "wb") as f:
            pickle.dump(index, f)

    def load_index_from_disk(self, filepath: str) -> PrecomputedIndex:
        """Reads files written by this class AND by the reference's indexer (same pickle layout, class path
        hilbert_quantization.core.precomputed_hilbert_index).  A restricted unpickler: only the two index dataclasses
        and NumPy's array reconstruction are resolvable, anything else in the stream raises."""
        level_cls, index_cls = self._level_cls or PrecomputedLevel, self._index_cls or PrecomputedIndex

        class _Restricted(pickle.Unpickler):
            def find_class(self, module, name):
                if module in (__name__, "hilbert_quantization.core.precomputed_hilbert_index"):
                    if name == "PrecomputedIndex":
                        return index_cls
                    if name == "PrecomputedLevel":
                        return level_cls
                if module.split(".")[0] == "numpy" and name in ("_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"):
                    return super().find_class(module, name)
                raise pickle.UnpicklingError(f"{module}.{name} is not allowed in a precomputed index file")

        with open(filepath, "rb") as f:
            index = _Restricted(f).load()
        self._index_cache[index.model_id] = index
        return index

    def get_storage_overhead(self, original_image_size: int) -> float:
        """core/precomputed_hilbert_index.py:233-262"""
        total = 0
        image_dim = int(np.sqrt(original_image_size // 4))
        for g, _ in self._calculate_granularity_levels(image_dim):
            total += (g * g + max(0, (g - 1) * (g - 1))) * (4 + 8)
        return (total / original_image_size) * 100
