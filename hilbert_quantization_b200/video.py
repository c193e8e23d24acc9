"""Video-path hierarchical similarity and frame ordering on the device.

core/video_search.py:215-262 (`_hierarchical_search`) with :1316-1328 scores every stored frame against the query
with the core engine's finest-level comparison: `hierarchical_search` below does that as ONE launch of
hq_core_level_sims over the whole frame pool.

core/video_storage.py:741-781 (`_traditional_search`, `_calculate_hierarchical_similarity`), :1203-1277
(`_sort_frames_by_hierarchical_indices`) and :1751-1803 (`_find_optimal_insertion_position`) call a NumPy
`corrcoef` per frame pair; here every similarity of a call is one launch of hq_pearson01_matrix (float64, like
np.corrcoef).  The greedy chain / insertion bookkeeping over the resulting matrix stays host code (it is
O(F^2) scalar comparisons on values that are already computed); the MPEG writer it feeds is the reference's.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _device as dev
from ._lib import check, lib


def hierarchical_similarity_matrix(a, b, device=None) -> torch.Tensor:
    """(pearson + 1) / 2 clamped to [0, 1] for all pairs of rows of a [M, S] and b [N, S] (common prefix S)."""
    d = dev.require_cuda(device)
    ta = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(d) if not isinstance(a, torch.Tensor) else a.to(d, torch.float64).contiguous()
    tb = torch.as_tensor(np.ascontiguousarray(np.asarray(b, dtype=np.float64))).to(d) if not isinstance(b, torch.Tensor) else b.to(d, torch.float64).contiguous()
    if ta.dim() == 1:
        ta = ta.reshape(1, -1)
    if tb.dim() == 1:
        tb = tb.reshape(1, -1)
    S = min(ta.shape[1], tb.shape[1])
    out = torch.empty((ta.shape[0], tb.shape[0]), dtype=torch.float64, device=d)
    if S == 0:
        return out.zero_()
    with torch.cuda.device(d):
        check(lib.hq_pearson01_matrix(dev.ptr(ta), ta.shape[0], ta.stride(0), dev.ptr(tb), tb.shape[0], tb.stride(0), S,
                                      dev.ptr(out), out.stride(0), dev.stream_ptr()))
    return out


def calculate_hierarchical_similarity(query_indices: np.ndarray, candidate_indices: np.ndarray, device=None) -> float:
    """core/video_storage.py:763-781"""
    if len(query_indices) == 0 or len(candidate_indices) == 0:
        return 0.0
    return float(hierarchical_similarity_matrix(query_indices, candidate_indices, device)[0, 0].item())


def traditional_search(query_indices: np.ndarray, frame_indices: Sequence[np.ndarray], max_results: int, device=None) -> List[Tuple[int, float]]:
    """core/video_storage.py:741-761: every stored frame scored against the query, stable descending sort,
    first `max_results` (frame position, similarity)."""
    if not len(frame_indices):
        return []
    sims = _row_similarities(query_indices, frame_indices, device)
    order = np.argsort(-sims, kind="stable")[:max_results]
    return [(int(i), float(sims[i])) for i in order]


def _row_similarities(query_indices, frame_indices, device) -> np.ndarray:
    """similarity of the query with each frame; frames may have different lengths (common prefix per pair)."""
    sims = np.zeros(len(frame_indices))
    q = np.asarray(query_indices, dtype=np.float64)
    by_len = {}
    for i, f in enumerate(frame_indices):
        by_len.setdefault(min(len(f), len(q)), []).append(i)
    for S, rows in by_len.items():
        if S == 0:
            continue
        block = np.stack([np.asarray(frame_indices[i], dtype=np.float64)[:S] for i in rows])
        sims[np.asarray(rows)] = hierarchical_similarity_matrix(q[:S], block, device)[0].cpu().numpy()
    return sims


def sort_frames_by_hierarchical_indices(frame_indices: Sequence[np.ndarray], device=None) -> List[int]:
    """core/video_storage.py:1203-1277: start at the frame closest to the centroid, then greedily append the
    remaining frame most similar to the last one (first wins on ties).  Returns the new order as positions
    into `frame_indices` (all index vectors of one video have the same length)."""
    F = len(frame_indices)
    if F <= 1:
        return list(range(F))
    lens = {len(f) for f in frame_indices}
    if len(lens) != 1 or 0 in lens:
        raise NotImplementedError("frame ordering is implemented for frames with index vectors of one common, non-zero length")
    X = np.stack([np.asarray(f, dtype=np.float64) for f in frame_indices])
    sim = hierarchical_similarity_matrix(X, X, device).cpu().numpy()
    centroid = np.mean([np.asarray(f) for f in frame_indices], axis=0)
    dist = np.array([np.linalg.norm(np.asarray(f) - centroid) for f in frame_indices])
    order = [int(np.argmin(dist))]                       # strict '<' in the reference: first minimum
    remaining = [i for i in range(F) if i != order[0]]
    while remaining:
        s = sim[order[-1], remaining]
        j = int(np.argmax(s)) if s.max() > -1.0 else 0   # strict '>' from -1.0: first maximum
        order.append(remaining.pop(j))
    return order


def find_optimal_insertion_position(new_indices: np.ndarray, existing_indices: Sequence[np.ndarray], device=None) -> int:
    """core/video_storage.py:1751-1803"""
    if not len(existing_indices):
        return 0
    sims = _row_similarities(new_indices, existing_indices, device)
    best_position, best_score = 0, -1.0
    if sims[0] > best_score:
        best_score, best_position = sims[0], 0
    for i in range(len(sims) - 1):
        score = (sims[i] + sims[i + 1]) / 2.0
        if score > best_score:
            best_score, best_position = score, i + 1
    if sims[-1] > best_score:
        best_score, best_position = sims[-1], len(sims)
    return int(best_position)


def hierarchical_search(query_indices: np.ndarray, frame_indices: Sequence, max_results: int,
                        similarity_threshold: float = 0.1, device=None) -> List[Tuple[int, float]]:
    """core/video_search.py:215-262 (`VideoEnhancedSearchEngine._hierarchical_search`): every frame that has index
    vectors is scored with `compare_indices_at_level(query, frame, 0)` (:1316-1328 -> core/search_engine.py:111-189),
    kept when the score is strictly above the threshold, stable descending sort, first `max_results`
    (frame position, similarity).  `None` entries are frames without index vectors (skipped, :237)."""
    from .search import ProgressiveSimilaritySearchEngine
    have = [i for i, f in enumerate(frame_indices) if f is not None]
    if not have:
        return []
    if len(query_indices) == 0:
        sims = np.zeros(len(have))                      # :1321-1322
    else:
        eng = ProgressiveSimilaritySearchEngine(similarity_threshold, device=device)
        all_levels = eng._level_sims(query_indices, [frame_indices[i] for i in have])
        sims = all_levels[:, 0] if all_levels.shape[1] else np.zeros(len(have))
    keep = np.nonzero(sims > similarity_threshold)[0]
    order = keep[np.argsort(-sims[keep], kind="stable")][:max_results]
    return [(have[int(j)], float(sims[j])) for j in order]
