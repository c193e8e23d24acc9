"""Hierarchical index generators (variants A, B, C) on the fused sm_100a tile kernel.

Interfaces mirrored (names, argument order, return types):
  * HierarchicalIndexGeneratorImpl      core/index_generator.py:13-356      (A, or B when
    config.use_streaming_optimization)  -- ABC interfaces.py:87-143
  * StreamingHilbertIndexGenerator      core/streaming_index_builder.py:274-343   (B)
  * HierarchicalIndexGenerator          rag/embedding_generation/hierarchical_index_generator.py:14  (C)
Batched entry point used by the database / benchmarks: `map_and_index`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _device as dev
from . import plans
from ._lib import check, lib
from .mapper import HilbertCurveMapper

_PLAN_CACHE: Dict[tuple, torch.Tensor] = {}


def _plan_tensor(key: tuple, plan: np.ndarray, device: torch.device) -> torch.Tensor:
    k = (key, str(device))
    t = _PLAN_CACHE.get(k)
    if t is None:
        t = torch.from_numpy(plan).to(device)
        _PLAN_CACHE[k] = t
    return t


def fused_pass(src: torch.Tensor, direction: int, n: int, D: int, *, plan: Optional[np.ndarray] = None,
               plan_key: tuple = (), min_level: int = 99, pyr_mode: int = 0, want_grid: bool = False,
               want_stream: bool = False, grid_out: Optional[torch.Tensor] = None,
               idx_out: Optional[torch.Tensor] = None, idx_stride: Optional[int] = None,
               grid_stride: Optional[int] = None):
    """One launch of hq_map_index_fused_ml.  src: float32 [N, D] (direction 0) or
    [N, n, n] (direction 1) on a CUDA device.  Returns (grid, stream, idx)."""
    d = dev.require_cuda(src.device)
    if src.dtype != torch.float32:
        raise TypeError("fused pass works on float32 items")
    src = src.contiguous()
    N = src.shape[0]
    cells = n * n
    src_stride = src.stride(0) if N > 1 else (D if direction == 0 else cells)      # (a size-1 axis may report any stride)
    grid = stream = idx = None
    if direction == 0 and (want_grid or grid_out is not None):
        grid = grid_out if grid_out is not None else torch.empty((N, n, n), dtype=torch.float32, device=d)
        g_stride = grid_stride if grid_stride is not None else cells
    else:
        g_stride = 0
    if direction == 1 and want_stream:
        stream = torch.empty((N, D), dtype=torch.float32, device=d)
    plan_t, plan_len = None, 0
    if plan is not None and len(plan) > 0:
        plan_t = _plan_tensor(plan_key, plan, d)
        plan_len = int(plan_t.numel())
        if idx_out is not None:
            idx = idx_out
        else:
            idx = torch.empty((N, plan_len), dtype=torch.float64 if pyr_mode else torch.float32, device=d)
        i_stride = idx_stride if idx_stride is not None else plan_len
    else:
        i_stride = 0
    scratch, scratch_bytes = None, 0
    if n > 64 and plan_len > 0:
        scratch_bytes = int(lib.hq_fused_scratch_bytes_min_level(N, n, pyr_mode, min(min_level, 6)))
        scratch = torch.empty(max(scratch_bytes, 8), dtype=torch.uint8, device=d)
    with torch.cuda.device(d):
        check(lib.hq_map_index_fused_ml(dev.ptr(src), direction, N, D, src_stride, n,
                                        dev.ptr(grid), g_stride, dev.ptr(stream), D if stream is not None else 0,
                                        dev.ptr(plan_t), plan_len, pyr_mode, min_level,
                                        dev.ptr(idx), i_stride, dev.ptr(scratch), scratch_bytes, dev.stream_ptr()))
    return grid, stream, idx


def map_and_index(embeddings: torch.Tensor, n: Optional[int] = None, *, variant: str = "C", index_space: Optional[int] = None,
                  layout: str = "compact", want_grid: bool = True, enhanced: bool = False):
    """Batched map_to_2d + hierarchical index in ONE kernel pass.

    embeddings: float32 [N, D] on the GPU.  Returns (grids or None, indices).
      variant "C": indices float32 [N, sum g^2] (layout 'compact') or [N, L, n] ('rows');
                   enhanced=True returns a single [N, n + L, n] frame tensor (grid + index rows),
                   the layout of generate_multi_level_indices.
      variant "A": float32 [N, S];  variant "B": float64 [N, S]  (S = index_space or n).
    """
    N, D = embeddings.shape
    if n is None:
        from .dimension import rag_optimal_dimensions
        n = rag_optimal_dimensions(D)[0]
    if variant == "C":
        plan, widths, ml = plans.c_plan(n, "rows" if enhanced else layout)
        L = len(widths)
        if enhanced:
            frames = torch.empty((N, n + L, n), dtype=torch.float32, device=embeddings.device)
            flat = frames.view(N, -1)
            fused_pass(embeddings, 0, n, D, plan=plan, plan_key=("C", n, "rows"), min_level=ml, grid_out=flat,
                       grid_stride=(n + L) * n, idx_out=flat[:, n * n:], idx_stride=(n + L) * n)
            return frames, None
        grid, _, idx = fused_pass(embeddings, 0, n, D, plan=plan, plan_key=("C", n, layout), min_level=ml, want_grid=want_grid)
        if layout == "rows":
            idx = idx.view(N, L, n)
        return grid, idx
    S = int(index_space if index_space is not None else n)
    if variant == "A":
        plan, ml = plans.a_plan(n, S)
        grid, _, idx = fused_pass(embeddings, 0, n, D, plan=plan, plan_key=("A", n, S), min_level=ml, want_grid=want_grid)
        return grid, idx
    if variant == "B":
        plan, ml = plans.b_plan(n, S)
        grid, _, idx = fused_pass(embeddings, 0, n, D, plan=plan, plan_key=("B", n, S), min_level=ml, pyr_mode=1, want_grid=want_grid)
        return grid, idx
    raise ValueError(f"unknown index variant {variant!r}")


def map_parameter_stream(stream: torch.Tensor, n: int = 4096, *, variant: str = "C", index_space: Optional[int] = None,
                         want_grid: bool = True, grid_out: Optional[torch.Tensor] = None,
                         idx_out: Optional[torch.Tensor] = None):
    """A flat float32 parameter stream (BASELINE config 4: a model's `named_parameters()` concatenated,
    core/streaming_processor.py:414-432,539-582) -> ceil(len / n^2) Hilbert grids of n x n with their hierarchical
    indices, the last grid zero padded (core/pipeline.py:325-349).  One launch over all grids when the stream is
    16-byte aligned and its length a multiple of 4 (hq_map_index_stream); otherwise the full grids and the tail grid go
    through two fused passes.  Returns (grids [N, n, n] or None, indices [N, S])."""
    d = dev.require_cuda(stream.device)
    if stream.dtype != torch.float32 or stream.dim() != 1:
        raise TypeError("parameter stream must be a 1-D float32 tensor")
    stream = stream.contiguous()
    total, cells = int(stream.numel()), n * n
    N = -(-total // cells)
    pyr_mode = 0
    if variant == "C":
        plan, _, ml = plans.c_plan(n, "compact")
        key = ("C", n, "compact")
    elif variant == "A":
        S = int(index_space if index_space is not None else n)
        plan, ml = plans.a_plan(n, S)
        key = ("A", n, S)
    elif variant == "B":
        S = int(index_space if index_space is not None else n)
        plan, ml = plans.b_plan(n, S)
        key, pyr_mode = ("B", n, S), 1
    else:
        raise ValueError(f"unknown index variant {variant!r}")
    grid = grid_out
    if grid is None and want_grid:
        grid = torch.empty((N, n, n), dtype=torch.float32, device=d)
    idx = idx_out if idx_out is not None else torch.empty((N, len(plan)), dtype=torch.float64 if pyr_mode else torch.float32, device=d)
    if N == 0:
        return grid, idx
    one_launch = n > 64 and total % 4 == 0 and stream.data_ptr() % 16 == 0 and (grid is None or grid.data_ptr() % 16 == 0) \
        and N * n < (1 << 31)
    if one_launch:
        plan_t = _plan_tensor(key, plan, d)
        sb = int(lib.hq_fused_scratch_bytes_min_level(N, n, pyr_mode, min(ml, 6)))
        scratch = torch.empty(max(sb, 8), dtype=torch.uint8, device=d)
        with torch.cuda.device(d):
            check(lib.hq_map_index_stream(dev.ptr(stream), total, n, dev.ptr(grid), dev.ptr(plan_t), int(plan_t.numel()), pyr_mode, ml,
                                          dev.ptr(idx), idx.stride(0), dev.ptr(scratch), sb, dev.stream_ptr()))
        return grid, idx
    full = total // cells
    if full:
        fused_pass(stream[: full * cells].view(full, cells), 0, n, cells, plan=plan, plan_key=key, min_level=ml, pyr_mode=pyr_mode,
                   grid_out=grid[:full].view(full, -1) if grid is not None else None, idx_out=idx[:full])
    if N > full:
        tail = total - full * cells
        fused_pass(stream[full * cells:].view(1, tail), 0, n, tail, plan=plan, plan_key=key, min_level=ml, pyr_mode=pyr_mode,
                   grid_out=grid[full:].view(1, -1) if grid is not None else None, idx_out=idx[full:])
    return grid, idx


def index_from_grids(grids: torch.Tensor, *, variant: str = "C", index_space: Optional[int] = None, layout: str = "compact"):
    """Hierarchical indices of already-mapped float32 grids [N, n, n] (direction 1 pass)."""
    N, n, _ = grids.shape
    if variant == "C":
        plan, widths, ml = plans.c_plan(n, layout)
        _, _, idx = fused_pass(grids, 1, n, n * n, plan=plan, plan_key=("C", n, layout), min_level=ml)
        return idx.view(N, len(widths), n) if layout == "rows" else idx
    S = int(index_space if index_space is not None else n)
    if variant == "A":
        plan, ml = plans.a_plan(n, S)
        return fused_pass(grids, 1, n, n * n, plan=plan, plan_key=("A", n, S), min_level=ml)[2]
    plan, ml = plans.b_plan(n, S)
    return fused_pass(grids, 1, n, n * n, plan=plan, plan_key=("B", n, S), min_level=ml, pyr_mode=1)[2]


def _square_pow2(image: np.ndarray) -> bool:
    return image.ndim == 2 and image.shape[0] == image.shape[1] and plans.is_pow2(image.shape[0])


def _block_means(image_t: torch.Tensor, sh: int, sw: int, rows: np.ndarray, cols: np.ndarray) -> torch.Tensor:
    """Generic rectangular block means of one float32 image on the device."""
    d = image_t.device
    H, W = image_t.shape
    r = torch.from_numpy(np.asarray(rows, dtype=np.int32)).to(d)
    c = torch.from_numpy(np.asarray(cols, dtype=np.int32)).to(d)
    out = torch.empty(len(rows), dtype=torch.float32, device=d)
    with torch.cuda.device(d):
        check(lib.hq_block_means(dev.ptr(image_t), 1, H, W, H * W, sh, sw, dev.ptr(r), dev.ptr(c), len(rows),
                                 dev.ptr(out), len(rows), dev.stream_ptr()))
    return out


class StreamingHilbertIndexGenerator:
    """Variant B (core/streaming_index_builder.py:274-343)."""

    def __init__(self, device=None):
        self.hilbert_mapper = HilbertCurveMapper(device)
        self._device = device

    def generate_optimized_indices(self, image: np.ndarray, index_space_size: int) -> np.ndarray:
        height, width = image.shape
        if width != height or width <= 0 or (width & (width - 1)) != 0:
            raise ValueError(f"Image must be square with power-of-2 dimensions, got {width}x{height}")
        if index_space_size <= 0:
            return np.array([])
        d = dev.require_cuda(self._device)
        g = dev.f32_device(image, d).reshape(1, width, width)
        if width < 4:
            # 1x1 / 2x2 images are below the fused kernel's 4x4 unit: the same plan gathered from a device vector of
            # [cells row-major | level-1 mean in the reference's ((a + b) + c) + d association] (float64, like the builder)
            plan, _ = plans.b_plan(width, index_space_size)
            cells = g.reshape(-1).to(torch.float64)
            pyr = cells
            if width == 2:
                s = cells[torch.from_numpy(plans.cells_of_positions(2, np.arange(4))).to(d)]
                pyr = torch.cat([cells, ((((s[0] + s[1]) + s[2]) + s[3]) * 0.25).reshape(1)])
            pt = torch.from_numpy(plan.astype(np.int64)).to(d)
            out = torch.where(pt >= 0, pyr[pt.clamp(min=0)], torch.zeros((), dtype=torch.float64, device=d))
            return out.cpu().numpy()
        return index_from_grids(g, variant="B", index_space=index_space_size)[0].cpu().numpy()

    def generate_indices_during_mapping(self, parameters: np.ndarray, dimensions: tuple, index_space_size: int) -> tuple:
        """Image + indices fed with the len(parameters) real values only
        (core/streaming_index_builder.py:287-313, core/hilbert_mapper.py:157-172)."""
        width, height = dimensions
        image = self.hilbert_mapper.map_to_2d(parameters, dimensions)
        count = len(parameters)
        # level k receives floor(count / 4^k) values; only complete groups of four promote
        sizes, c = [], count
        while len(sizes) < plans.B_MAX_LEVELS and c >= 1:
            sizes.append(c)
            c //= 4
        stats = {"total_values_processed": count, "levels_used": len(sizes),
                 "indices_per_level": {k: s for k, s in enumerate(sizes)},
                 "current_window_sizes": {k: (s % 4) for k, s in enumerate(sizes)},
                 "level_counters": sizes + [0] * (plans.B_MAX_LEVELS - len(sizes))}
        if index_space_size <= 0:
            return image, np.array([]), stats
        d = dev.require_cuda(self._device)
        alloc = plans.b_allocations(sizes, index_space_size)
        entries: List[int] = []
        n = width
        for k, a in enumerate(alloc):
            if a <= 0 or k >= len(sizes):
                continue
            sz = sizes[k]
            pos = np.array([int(i * (sz / a)) for i in range(a)], dtype=np.int64) if sz > a else np.arange(sz, dtype=np.int64)
            entries.extend((plans.cells_of_positions(n, pos) if k == 0 else plans.level_base(n, k) + pos).tolist())
        plan = np.full(index_space_size, -1, dtype=np.int64)
        m = min(len(entries), index_space_size)
        plan[:m] = entries[:m]
        plan = plan.astype(np.int32)
        src = dev.f32_device(np.asarray(parameters).reshape(1, -1), d)
        _, _, idx = fused_pass(src, 0, n, count, plan=plan, plan_key=("Bdm", n, index_space_size, count),
                               min_level=plans._min_level(n, plan), pyr_mode=1)
        return image, idx[0].cpu().numpy(), stats


class HierarchicalIndexGeneratorImpl:
    """Variant A, or B when config.use_streaming_optimization (core/index_generator.py:13-356)."""

    def __init__(self, config=None, device=None):
        self.config = config
        self._device = device
        self._streaming_generator = (StreamingHilbertIndexGenerator(device)
                                     if getattr(config, "use_streaming_optimization", False) else None)

    # -- host-side scalar logic (identical results to the reference) --
    def calculate_level_allocation(self, total_space: int) -> List[Tuple[int, int]]:
        return plans.a_allocation(total_space)

    def calculate_spatial_averages(self, image: np.ndarray, grid_size: int) -> List[float]:
        if image.size == 0 or grid_size <= 0:
            return []
        h, w = image.shape
        sh, sw = h // grid_size, w // grid_size
        d = dev.require_cuda(self._device)
        img = dev.f32_device(image, d)
        if sh == 0 or sw == 0:
            return [float(_block_means(img, h, w, np.zeros(1), np.zeros(1))[0].item())]
        rows, cols = np.divmod(np.arange(grid_size * grid_size), grid_size)
        return [float(v) for v in _block_means(img, sh, sw, rows, cols).cpu().tolist()]

    def calculate_offset_samples(self, image: np.ndarray, section_size: int, available_space: int) -> List[float]:
        if image.size == 0:
            return []
        h, w = image.shape
        pos = plans.a_sample_positions(h, w, section_size, available_space)
        return [float(image[r, c]) for r, c in pos]

    def embed_indices_in_image(self, image: np.ndarray, indices: np.ndarray) -> np.ndarray:
        """core/index_generator.py:221-253 (host copy; one extra row)."""
        if image.size == 0:
            return image
        h, w = image.shape
        out = np.zeros((h + 1, w), dtype=image.dtype)
        out[:h] = image
        m = min(len(indices), w)
        out[h, :m] = indices[:m]
        return out

    def extract_indices_from_image(self, enhanced_image: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """core/index_generator.py:255-290."""
        if enhanced_image.size == 0:
            return enhanced_image, np.array([])
        h, w = enhanced_image.shape
        if h < 2:
            return enhanced_image, np.array([])
        row = enhanced_image[-1, :]
        nz = np.nonzero(row)[0]
        row = row[: nz[-1] + 1] if len(nz) > 0 else (row[:1] if len(row) > 0 else np.array([]))
        return enhanced_image[:-1, :], row

    def generate_optimized_indices(self, image: np.ndarray, index_space_size: int) -> np.ndarray:
        if image.size == 0 or index_space_size <= 0:
            return np.array([])
        if self._streaming_generator is not None:
            return self._streaming_generator.generate_optimized_indices(image, index_space_size)
        return self._generate_traditional_indices(image, index_space_size)

    def generate_indices_with_integrated_mapping(self, parameters: np.ndarray, dimensions: Tuple[int, int],
                                                 index_space_size: int) -> Tuple[np.ndarray, np.ndarray]:
        """The hook QuantizationPipeline.quantize_model probes for (core/pipeline.py:115-121; no reference class
        implements it): map_to_2d + hierarchical indices of one parameter vector in ONE fused launch.  Same results as
        `hilbert_mapper.map_to_2d` followed by `generate_optimized_indices` (bit-identical image; variant B bit-identical,
        variant A from the same kernel)."""
        width, height = dimensions
        params = np.asarray(parameters)
        if width != height or not plans.is_pow2(width) or params.ndim != 1 or len(params) > width * height:
            image = HilbertCurveMapper(self._device).map_to_2d(params, dimensions)        # raises the reference's errors
            return image, self.generate_optimized_indices(image, index_space_size)
        if params.dtype != np.float32 or width < 4 or index_space_size <= 0 or len(params) == 0:
            image = HilbertCurveMapper(self._device).map_to_2d(params, dimensions)
            return image, self.generate_optimized_indices(image, index_space_size)
        d = dev.require_cuda(self._device)
        src = dev.f32_device(params.reshape(1, -1), d)
        variant = "B" if self._streaming_generator is not None else "A"
        grid, idx = map_and_index(src, width, variant=variant, index_space=index_space_size, want_grid=True)
        return grid[0].cpu().numpy(), idx[0].cpu().numpy()

    def _generate_traditional_indices(self, image: np.ndarray, index_space_size: int) -> np.ndarray:
        d = dev.require_cuda(self._device)
        if _square_pow2(image) and image.shape[0] >= 4:
            g = dev.f32_device(image, d).reshape(1, image.shape[0], image.shape[0])
            return index_from_grids(g, variant="A", index_space=index_space_size)[0].cpu().numpy()
        # general rectangles: block means by the generic kernel, samples are plain reads
        h, w = image.shape
        img = dev.f32_device(image, d)
        vals: List[float] = []
        for seg in plans.a_segments(h, w, index_space_size):
            if seg[0] == "sample":
                vals.extend(float(image[r, c]) for r, c in seg[1])
            else:
                _, g, take = seg
                vals.extend(self.calculate_spatial_averages(image, g)[:take])
        res = np.zeros(index_space_size, dtype=np.float32)
        m = min(len(vals), index_space_size)
        res[:m] = np.asarray(vals[:m], dtype=np.float32)
        return res


class HierarchicalIndexGenerator:
    """Variant C (rag/embedding_generation/hierarchical_index_generator.py:14-627)."""

    def __init__(self, config=None, device=None):
        self.config = config or {}
        get = self.config.get if hasattr(self.config, "get") else (lambda k, dflt: getattr(self.config, k, dflt))
        self.min_granularity = get("min_granularity", 2)
        self.max_index_rows = get("max_index_rows", 8)
        self._device = device

    # -- scalar planning, same results as the reference --
    def _nearest_power_of_2(self, n: int) -> int:
        if n <= 0:
            return 1
        p = 1
        while p * 2 <= n:
            p *= 2
        return p

    def calculate_optimal_granularity(self, image_dimensions: Tuple[int, int]) -> Dict[str, object]:
        width, height = image_dimensions
        levels = plans.c_levels(width, self.min_granularity, self.max_index_rows)
        finest = self._nearest_power_of_2(max(self.min_granularity, int(math.sqrt(width))))
        return {"finest_granularity": finest, "granularity_levels": levels, "index_rows_needed": len(levels),
                "total_image_height": height + len(levels), "original_dimensions": image_dimensions,
                "section_sizes": [(width // g, height // g) for g in levels]}

    def allocate_index_space(self, image_dimensions: Tuple[int, int]) -> Dict[str, object]:
        info = self.calculate_optimal_granularity(image_dimensions)
        width, height = image_dimensions
        rows = info["index_rows_needed"]
        return {"enhanced_dimensions": (width, height + rows),
                "index_row_positions": [height + i for i in range(rows)], "granularity_info": info}

    def _generate_hilbert_coordinates(self, n: int) -> List[Tuple[int, int]]:
        """(row, col) visiting order of the sections (hierarchical_index_generator.py:286-342)."""
        if n == 1:
            return [(0, 0)]
        if n == 2:
            return [(0, 0), (0, 1), (1, 1), (1, 0)]
        if n & (n - 1):
            n = self._nearest_power_of_2(n)
        x, y = plans._d2xy(n, np.arange(n * n))
        return list(zip(y.tolist(), x.tolist()))

    def _default_levels(self, width: int) -> List[int]:
        return plans.c_levels(width, self.min_granularity, self.max_index_rows)

    # -- device work --
    def _calculate_hilbert_order_averages(self, image: np.ndarray, granularity: int) -> np.ndarray:
        h, w = image.shape
        sh, sw = h // granularity, w // granularity
        d = dev.require_cuda(self._device)
        img = dev.f32_device(image, d)
        if sh == 0 or sw == 0:
            return _block_means(img, h, w, np.zeros(1), np.zeros(1)).cpu().numpy().astype(image.dtype)
        order = self._generate_hilbert_coordinates(granularity)
        rows = np.array([r for r, _ in order])
        cols = np.array([c for _, c in order])
        return _block_means(img, sh, sw, rows, cols).cpu().numpy().astype(image.dtype)

    def _calculate_spatial_averages(self, image: np.ndarray, granularity: int) -> np.ndarray:
        h, w = image.shape
        sh, sw = h // granularity, w // granularity
        d = dev.require_cuda(self._device)
        img = dev.f32_device(image, d)
        if sh == 0 or sw == 0:
            return _block_means(img, h, w, np.zeros(1), np.zeros(1)).cpu().numpy().astype(image.dtype)
        rows, cols = np.divmod(np.arange(granularity * granularity), granularity)
        return _block_means(img, sh, sw, rows, cols).cpu().numpy().astype(image.dtype)

    def create_progressive_granularity_levels(self, embedding_image: np.ndarray) -> List[np.ndarray]:
        if embedding_image.ndim != 2:
            raise ValueError("Embedding image must be 2D")
        h, w = embedding_image.shape
        levels = self._default_levels(w)
        if _square_pow2(embedding_image) and w >= 4 and levels == plans.c_levels(w):
            d = dev.require_cuda(self._device)
            g = dev.f32_device(embedding_image, d).reshape(1, w, w)
            compact = index_from_grids(g, variant="C", layout="compact")[0].cpu().numpy().astype(embedding_image.dtype)
            rows, o = [], 0
            for wd in plans.c_plan(w, "compact")[1]:
                rows.append(compact[o:o + wd])
                o += wd
            return rows
        return [self._calculate_hilbert_order_averages(embedding_image, g) for g in levels]

    def calculate_averages_for_multiple_granularities(self, embedding_image: np.ndarray,
                                                      granularity_levels: List[int]) -> Dict[int, np.ndarray]:
        if embedding_image.ndim != 2:
            raise ValueError("Embedding image must be 2D")
        return {g: self._calculate_hilbert_order_averages(embedding_image, g) for g in granularity_levels if g > 0}

    def generate_multi_level_indices(self, embedding_image: np.ndarray) -> np.ndarray:
        if embedding_image.ndim != 2:
            raise ValueError("Embedding image must be 2D")
        return self.embed_multi_level_indices(embedding_image, self.create_progressive_granularity_levels(embedding_image)) \
            if self._default_levels(embedding_image.shape[1]) else embedding_image.copy()

    def create_enhanced_embedding_with_indices(self, embedding_image: np.ndarray) -> np.ndarray:
        if embedding_image.ndim != 2:
            raise ValueError("Embedding image must be 2D")
        return self.embed_multi_level_indices(embedding_image, self.create_progressive_granularity_levels(embedding_image))

    def embed_multi_level_indices(self, image: np.ndarray, index_rows: List[np.ndarray]) -> np.ndarray:
        """hierarchical_index_generator.py:344-385 (host copy)."""
        if image.ndim != 2:
            raise ValueError("Image must be 2D")
        if not index_rows:
            return image.copy()
        h, w = image.shape
        out = np.zeros((h + len(index_rows), w), dtype=image.dtype)
        out[:h] = image
        for i, row in enumerate(index_rows):
            if len(row) <= w:
                out[h + i, : len(row)] = row
            else:
                out[h + i, :] = row[:w]
        return out

    def extract_indices_from_image(self, enhanced_image: np.ndarray, original_height: int = None):
        """hierarchical_index_generator.py:387-441.  Without `original_height` the reference
        guesses it heuristically (:443-506); the device path always carries
        the height explicitly, so here a missing hint assumes the default number of index rows."""
        if enhanced_image.ndim != 2:
            raise ValueError("Enhanced image must be 2D")
        height, width = enhanced_image.shape
        if original_height is None:
            original_height = self._detect_original_image_height(enhanced_image)
        original_height = max(0, min(original_height, height))
        if original_height >= height:
            return enhanced_image, []
        rows = []
        for r in range(original_height, height):
            row = enhanced_image[r, :]
            nz = np.nonzero(row)[0]
            rows.append(row[: nz[-1] + 1] if len(nz) > 0 else (row[:1] if len(row) > 0 else np.array([])))
        return enhanced_image[:original_height, :], rows

    def _detect_original_image_height(self, enhanced_image: np.ndarray) -> int:
        """Explicit layout: the frame carries the default number of index rows below the grid.  (The reference guesses the
        height from row sparsity and variances, hierarchical_index_generator.py:443-506, SURVEY 9.6 quirk 1;
        `dropin.install()` keeps the reference's guess for this one method so unchanged callers see its behaviour.)"""
        height, width = enhanced_image.shape
        return height - len(self._default_levels(width))

    def validate_index_allocation(self, image_dimensions: Tuple[int, int]) -> bool:
        try:
            levels = self.allocate_index_space(image_dimensions)["granularity_info"]["granularity_levels"]
            width, height = image_dimensions
            return bool(levels) and levels[0] <= min(width, height) and len(levels) <= self.max_index_rows
        except Exception:
            return False
