"""Host-side planning: which pyramid entries make up each index variant.

Every index variant of the reference is "run means of the Hilbert-ordered stream plus a
fixed gather" (the gather only depends on the grid side and the index size), so the host
computes the gather once as an int32 plan and the fused kernel applies it per item.
Plan encoding (include/hq_b200.h): -1 -> 0.0; [0, n*n) -> grid cell (row-major);
n*n + level_base(k) + j -> run mean j of level k (runs of 4^k curve positions).

Only index arithmetic happens here (no embedding data is touched on the host).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Tuple

import numpy as np

B_MAX_LEVELS = 10        # core/streaming_index_builder.py:21


def is_pow2(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def level_base(n: int, k: int) -> int:
    """Offset of level k (>= 1) inside the plan's pyramid address space."""
    cells = n * n
    return cells + sum(cells >> (2 * i) for i in range(1, k))


def _d2xy(n: int, d: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Curve position -> (x, y) for plan construction (core/hilbert_mapper.py:42-66)."""
    t = np.asarray(d, dtype=np.int64).copy()
    x = np.zeros_like(t)
    y = np.zeros_like(t)
    s = 1
    while s < n:
        rx = (t >> 1) & 1
        ry = (t ^ rx) & 1
        turn = ry == 0
        mirror = turn & (rx == 1)
        x = np.where(mirror, s - 1 - x, x)
        y = np.where(mirror, s - 1 - y, y)
        x, y = np.where(turn, y, x), np.where(turn, x, y)
        x += s * rx
        y += s * ry
        t >>= 2
        s <<= 1
    return x, y


def _xy2d(n: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """(x, y) -> curve position (core/hilbert_mapper.py:68-90)."""
    x = np.asarray(x, dtype=np.int64).copy()
    y = np.asarray(y, dtype=np.int64).copy()
    d = np.zeros_like(x)
    s = n >> 1
    while s > 0:
        rx = ((x & s) != 0).astype(np.int64)
        ry = ((y & s) != 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        turn = ry == 0
        mirror = turn & (rx == 1)
        x = np.where(mirror, s - 1 - x, x)
        y = np.where(mirror, s - 1 - y, y)
        x, y = np.where(turn, y, x), np.where(turn, x, y)
        s >>= 1
    return d


def cells_of_positions(n: int, d: np.ndarray) -> np.ndarray:
    x, y = _d2xy(n, d)
    return y * n + x


# ---------------------------------------------------------------- variant C (RAG)
def c_levels(width: int, min_granularity: int = 2, max_index_rows: int = 8) -> List[int]:
    """rag/embedding_generation/hierarchical_index_generator.py:23-68."""
    g = max(min_granularity, int(math.sqrt(width)))
    p = 1
    while p * 2 <= g:
        p *= 2
    out = []
    while p >= min_granularity and len(out) < max_index_rows:
        out.append(p)
        p //= 2
    return out


def _c_level_entries(n: int, g: int) -> np.ndarray:
    """Plan entries of the row of granularity g on an n x n grid (sections in the
    reference's visiting order, hierarchical_index_generator.py:204-244, :286-342)."""
    k = int(math.log2(n // g))                    # block = (n/g)^2 = 4^k cells
    if g == 2:
        order = np.array([0, 3, 2, 1])            # hard-coded [(0,0),(0,1),(1,1),(1,0)] as (row, col)
    else:
        order = np.arange(g * g)
    if k == 0:                                    # one cell per section
        if g == 2:
            return np.array([0, 1, n + 1, n], dtype=np.int64)
        return cells_of_positions(n, order)
    return level_base(n, k) + order


@lru_cache(maxsize=None)
def c_plan(n: int, layout: str = "compact") -> Tuple[np.ndarray, List[int], int]:
    """(plan, level widths, min level).  layout 'compact' = levels concatenated;
    'rows' = one zero-padded row of n values per level (the enhanced-frame layout)."""
    levels = c_levels(n)
    parts, widths = [], []
    for g in levels:
        e = _c_level_entries(n, g)[:n]
        widths.append(len(e))
        if layout == "rows":
            row = np.full(n, -1, dtype=np.int64)
            row[: len(e)] = e
            parts.append(row)
        else:
            parts.append(e)
    plan = np.concatenate(parts).astype(np.int32)
    return plan, widths, _min_level(n, plan)


def _min_level(n: int, plan: np.ndarray) -> int:
    pyr = plan[plan >= n * n]
    if pyr.size == 0:
        return 99
    off = int(pyr.min()) - n * n
    k = 1
    while off >= (n * n) >> (2 * k):
        off -= (n * n) >> (2 * k)
        k += 1
    return k


# ---------------------------------------------------------------- variant B (streaming)
def b_allocations(level_sizes: List[int], total_space: int) -> List[int]:
    """core/streaming_index_builder.py:207-243."""
    alloc = [0] * B_MAX_LEVELS
    live = [k for k in range(B_MAX_LEVELS) if k < len(level_sizes) and level_sizes[k] > 0]
    left = total_space
    for i, k in enumerate(live):
        if i == len(live) - 1:
            alloc[k] = left
        else:
            a = min(max(1, int(total_space * (0.5 ** (i + 1)))), left)
            alloc[k] = a
            left -= a
    return alloc


@lru_cache(maxsize=None)
def b_plan(n: int, index_space_size: int) -> Tuple[np.ndarray, int]:
    """Plan of StreamingHilbertIndexGenerator.generate_optimized_indices on a full n x n
    image (core/streaming_index_builder.py:154-205, :315-343)."""
    cells = n * n
    sizes = []
    c = cells
    while len(sizes) < B_MAX_LEVELS and c >= 1:
        sizes.append(c)
        if c < 4:
            break
        c //= 4
    alloc = b_allocations(sizes, index_space_size)
    entries: List[int] = []
    for k, a in enumerate(alloc):
        if a <= 0 or k >= len(sizes):
            continue
        sz = sizes[k]
        if sz > a:
            step = sz / a
            pos = np.array([int(i * step) for i in range(a)], dtype=np.int64)
        else:
            pos = np.arange(sz, dtype=np.int64)
        if k == 0:
            entries.extend(cells_of_positions(n, pos).tolist())
        else:
            entries.extend((level_base(n, k) + pos).tolist())
    plan = np.full(index_space_size, -1, dtype=np.int64)
    m = min(len(entries), index_space_size)
    plan[:m] = entries[:m]
    plan = plan.astype(np.int32)
    return plan, _min_level(n, plan)


# ---------------------------------------------------------------- variant A (traditional)
def a_allocation(total_space: int) -> List[Tuple[int, int]]:
    """core/index_generator.py:34-98."""
    if total_space <= 0:
        return []
    out = []
    left = total_space
    limit = min(32, int(math.sqrt(total_space)))
    g = 1
    while g <= limit:
        g *= 2
    g = max(g // 2, 2)
    frac = 0.5
    while left > 0 and g >= 1:
        take = min(int(left * frac), g * g, left)
        if take > 0:
            out.append((g, take))
            left -= take
        g //= 2
        frac *= 0.5
        if frac < 0.01:
            break
    if left > 0 and out:
        out.append((out[0][0], left))
    return out


def a_sample_positions(h: int, w: int, section: int, available: int) -> List[Tuple[int, int]]:
    """core/index_generator.py:146-219: 4 corners + centre per section, row-major."""
    if h * w == 0 or section <= 0 or available <= 0:
        return []
    sy, sx = h // section, w // section
    if sy == 0 or sx == 0:
        return [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, w // 2)][:available]
    todo = min(available // 5, sy * sx)
    pos: List[Tuple[int, int]] = []
    done = 0
    for r in range(sy):
        for c in range(sx):
            if done >= todo:
                break
            r0, r1 = r * section, min((r + 1) * section, h)
            c0, c1 = c * section, min((c + 1) * section, w)
            pos.extend([(r0, c0), (r0, c1 - 1), (r1 - 1, c0), (r1 - 1, c1 - 1), ((r0 + r1) // 2, (c0 + c1) // 2)])
            done += 1
            if len(pos) >= available:
                break
        if len(pos) >= available:
            break
    return pos[:available]


def a_segments(h: int, w: int, index_space_size: int):
    """Routing of _generate_traditional_indices (core/index_generator.py:313-356): the
    first emitted allocation is row-major block means, every later one whose grid size
    also appears in allocations[:-1] is offset sampling (:329-332)."""
    allocs = a_allocation(index_space_size)
    segs, emitted = [], 0
    for g, space in allocs:
        if space <= 0:
            continue
        if emitted > 0 and any(pg == g for pg, _ in allocs[:-1]):
            pos = a_sample_positions(h, w, max(1, h // g), space)
            segs.append(("sample", pos))
            emitted += len(pos)
        else:
            count = 1 if (h // g == 0 or w // g == 0) else g * g
            take = min(count, space)
            segs.append(("mean", g, take))
            emitted += take
    return segs


@lru_cache(maxsize=None)
def a_plan(n: int, index_space_size: int) -> Tuple[np.ndarray, int]:
    """Plan of variant A on a square power-of-two image."""
    entries: List[int] = []
    top = int(math.log2(n))
    for seg in a_segments(n, n, index_space_size):
        if seg[0] == "sample":
            entries.extend(r * n + c for r, c in seg[1])
            continue
        _, g, take = seg
        if g > n:                                  # grid finer than the image: overall mean
            entries.append(level_base(n, top) if top >= 1 else 0)
            continue
        k = int(math.log2(n // g))
        rows, cols = np.divmod(np.arange(g * g), g)
        if k == 0:
            e = rows * n + cols
        else:
            e = level_base(n, k) + _xy2d(g, cols, rows)
        entries.extend(e[:take].tolist())
    plan = np.full(index_space_size, -1, dtype=np.int64)
    m = min(len(entries), index_space_size)
    plan[:m] = entries[:m]
    plan = plan.astype(np.int32)
    return plan, _min_level(n, plan)


# ---------------------------------------------------------------- core search layout
def core_levels(length: int, total_space: int) -> List[Tuple[int, int, int, bool]]:
    """core/search_engine.py:42-109 -> [(grid, start, end, is_offset)]."""
    if length == 0 or total_space <= 0:
        return []
    out = []
    left, cur = total_space, 0
    limit = min(32, int(math.sqrt(total_space)))
    g = 1
    while g <= limit:
        g *= 2
    g = max(g // 2, 2)
    frac = 0.5
    seen = set()
    while left > 0 and g >= 1 and cur < length:
        take = min(int(left * frac), g * g, left)
        if take > 0:
            out.append((g, cur, cur + take, g in seen))
            seen.add(g)
            cur += take
            left -= take
        g //= 2
        frac *= 0.5
        if frac < 0.01:
            break
    if left > 0 and cur < length and out:
        out.append((out[0][0], cur, min(cur + left, length), True))
    return out
