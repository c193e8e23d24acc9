"""Batch containers either side of the hot path: the reference's storage format -- the ENHANCED FRAME,
an n x n Hilbert grid with one zero-padded index row per level appended below it -- as the device-resident database format,
with the original height carried explicitly instead of guessed.

  * `EmbeddingFrame`            host dataclass with the reference's fields (rag/models.py:38-59)
  * `EmbeddingFrameBatch`       N enhanced frames [N, H + L, W] float32 on the device + (H, D): built from embeddings by ONE
                                fused kernel (map_to_2d + index rows + embed, rag/.../hierarchical_index_generator.py:344-385),
                                or from stored frames; gives back the grids, the embeddings (map_from_2d), the compact index
                                rows the search uses (no recomputation) and reference-style `EmbeddingFrame` objects whose
                                `hierarchical_indices` are trimmed like the reference's extraction (:387-441: trailing zeros
                                stripped, all-zero rows dropped)
  * `EmbeddingDatabase.from_frames` (search.py) builds the search shard from such a batch
  * `QuantizedModelBatch`       the core path's candidates (models.py:55-79): `hierarchical_indices` of N QuantizedModel objects
                                stacked ONCE on the device, accepted by ProgressiveSimilaritySearchEngine as a candidate pool

The reference finds the original height of a stored frame with a ">= 50 % zeros" scan (rag/search/engine.py:134-162,
SURVEY 9.6 quirk 1), which mis-sizes padded frames; here the height is a field.  `strict=True` of RAGSearchEngineImpl keeps
the heuristic for parity tests."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _device as dev
from . import plans
from .dimension import rag_optimal_dimensions


@dataclass
class EmbeddingFrame:
    """rag/models.py:38-59 (same fields, same validation)."""
    embedding_data: np.ndarray
    hierarchical_indices: List[np.ndarray]
    original_embedding_dimensions: int
    hilbert_dimensions: Tuple[int, int]
    compression_quality: float
    frame_number: int

    def __post_init__(self):
        if self.original_embedding_dimensions <= 0:
            raise ValueError("Original embedding dimensions must be positive")
        if len(self.hilbert_dimensions) != 2:
            raise ValueError("Hilbert dimensions must be a 2-tuple")
        if self.compression_quality < 0 or self.compression_quality > 1:
            raise ValueError("Compression quality must be between 0 and 1")
        if self.frame_number < 0:
            raise ValueError("Frame number must be non-negative")
        if self.embedding_data.ndim != 2:
            raise ValueError("Embedding data must be 2-dimensional")


class EmbeddingFrameBatch:
    """N enhanced frames on the device.  `frames[:, :H]` are the grids, `frames[:, H + l, :]` the index row of level l
    (finest first, zero padded to the frame width)."""

    def __init__(self, frames: torch.Tensor, original_height: int, original_embedding_dimensions: int):
        if frames.dim() != 3 or frames.dtype != torch.float32 or not frames.is_cuda:
            raise TypeError("frames must be a float32 CUDA tensor [N, H + L, W]")
        H, W = int(original_height), int(frames.shape[2])
        if H != W or H & (H - 1):
            raise ValueError("the grid of an enhanced frame is a power-of-two square (original_height == width)")
        self.levels = plans.c_levels(W)
        if frames.shape[1] != H + len(self.levels):
            raise ValueError(f"a {W}-wide frame carries {len(self.levels)} index rows: expected height {H + len(self.levels)}, "
                             f"got {frames.shape[1]}")
        if not 0 < int(original_embedding_dimensions) <= H * W:
            raise ValueError("original_embedding_dimensions must be in (0, H * W]")
        self.frames = frames.contiguous()
        self.original_height = H
        self.original_embedding_dimensions = int(original_embedding_dimensions)

    # ---- constructors --------------------------------------------------------------------------------------------
    @classmethod
    def from_embeddings(cls, embeddings, n: Optional[int] = None, device=None) -> "EmbeddingFrameBatch":
        """[N, D] embeddings -> enhanced frames in one launch (hq_map_index with the frame as its grid + index output)."""
        from .index import map_and_index
        d = dev.require_cuda(device if device is not None else (embeddings.device if isinstance(embeddings, torch.Tensor)
                                                                 and embeddings.is_cuda else None))
        emb = dev.f32_device(embeddings, d)
        if emb.dim() != 2:
            raise ValueError("embeddings must be [N, D]")
        n = int(n) if n is not None else rag_optimal_dimensions(int(emb.shape[1]))[0]
        frames, _ = map_and_index(emb, n, variant="C", enhanced=True)
        return cls(frames, n, int(emb.shape[1]))

    @classmethod
    def from_frames(cls, frames, original_height: Optional[int] = None, original_embedding_dimensions: Optional[int] = None,
                    device=None) -> "EmbeddingFrameBatch":
        """Stored frames: a [N, H + L, W] array / tensor, or a sequence of EmbeddingFrame-like objects (`embedding_data`,
        `original_embedding_dimensions`; the reference's dataclass works).  The height is taken from the argument, else from
        the frame shape (height - index rows of that width) -- never from the pixel values."""
        d = dev.require_cuda(device)
        if not isinstance(frames, (np.ndarray, torch.Tensor)):
            objs = list(frames)
            if not objs:
                raise ValueError("no frames")
            if original_embedding_dimensions is None:
                dims = {int(o.original_embedding_dimensions) for o in objs}
                if len(dims) != 1:
                    raise ValueError("frames of one batch share their embedding dimension")
                original_embedding_dimensions = dims.pop()
            frames = np.stack([np.asarray(o.embedding_data, dtype=np.float32) for o in objs])
        t = dev.f32_device(frames, d)
        if t.dim() != 3:
            raise ValueError("frames must be [N, H + L, W]")
        W = int(t.shape[2])
        H = int(original_height) if original_height is not None else int(t.shape[1]) - len(plans.c_levels(W))
        D = int(original_embedding_dimensions) if original_embedding_dimensions is not None else H * W
        return cls(t, H, D)

    # ---- views ---------------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.frames.shape[0])

    @property
    def hilbert_dimensions(self) -> Tuple[int, int]:
        return (self.original_height, int(self.frames.shape[2]))

    def grids(self) -> torch.Tensor:
        return self.frames[:, : self.original_height, :]

    def embeddings(self) -> torch.Tensor:
        """[N, D]: map_from_2d of the grids (rag/.../hilbert_mapper.py:77-120), padding cells dropped."""
        from .mapper import HilbertCurveMapper
        return HilbertCurveMapper().map_from_2d_batch(self.grids().contiguous(), length=self.original_embedding_dimensions)

    def index_rows(self) -> torch.Tensor:
        """The compact index rows [N, Lsum] of the search (levels concatenated finest first, search.make_layout): sliced out
        of the frames, not recomputed."""
        H, W = self.original_height, int(self.frames.shape[2])
        return torch.cat([self.frames[:, H + l, : min(g * g, W)] for l, g in enumerate(self.levels)], dim=1).contiguous()

    def to_frames(self, compression_quality: float = 0.8, first_frame_number: int = 0) -> List[EmbeddingFrame]:
        """Host objects with the reference's fields; `hierarchical_indices` as the reference's extraction returns them."""
        host = self.frames.cpu().numpy()
        H = self.original_height
        out = []
        for i in range(host.shape[0]):
            rows = []
            for r in range(H, host.shape[1]):
                row = host[i, r, :]
                nz = np.nonzero(row)[0]
                if nz.size:                                   # hierarchical_index_generator.py:425-436
                    rows.append(row[: nz[-1] + 1].copy())
            out.append(EmbeddingFrame(host[i], rows, self.original_embedding_dimensions, self.hilbert_dimensions,
                                      float(compression_quality), first_frame_number + i))
        return out

    def database(self, **kw):
        from .search import EmbeddingDatabase
        return EmbeddingDatabase.from_frames(self, **kw)


class QuantizedModelBatch(Sequence):
    """A candidate pool of the core engine whose `hierarchical_indices` live on the device.

    Behaves like the list of QuantizedModel objects it was built from (len, indexing, iteration: the search results refer
    to the original objects) and carries, per distinct index length, the stacked float64 indices
    ProgressiveSimilaritySearchEngine compares against -- uploaded once instead of once per query."""

    def __init__(self, models: Sequence, device=None):
        self.models = list(models)
        d = dev.require_cuda(device)
        self.device = d
        by_len: Dict[int, List[int]] = {}
        for i, m in enumerate(self.models):
            by_len.setdefault(len(m.hierarchical_indices), []).append(i)
        self.groups: Dict[int, Tuple[np.ndarray, torch.Tensor]] = {}
        for S, rows in by_len.items():
            if S == 0:
                continue
            stack = np.stack([np.asarray(self.models[i].hierarchical_indices, dtype=np.float64) for i in rows])
            self.groups[S] = (np.asarray(rows), torch.from_numpy(stack).to(d))

    def __len__(self) -> int:
        return len(self.models)

    def __getitem__(self, i):
        return self.models[i]

    def __iter__(self):
        return iter(self.models)
