"""TEST INFRASTRUCTURE ONLY -- vectorised NumPy restatement of the reference hot path.

Reference: Tylerlhess/hilbert-quantization v1.3.0 (pure Python + NumPy).  All
``file:line`` citations are relative to the reference checkout
(``/root/reference`` in the authoring container).

Parity status: PINNED -- see ``oracle/pin_against_reference.py`` (function by
function against the reference's classes) and ``tests/golden/*.npz``
(reference outputs frozen by ``tests/golden/make_golden.py``).

Conventions
-----------
* "stream" = the 1-D Hilbert-ordered sequence; ``grid[y, x]`` with
  ``(x, y) = d2xy(n, d)`` holds ``stream[d]``.
* Integer / index / byte results are bit-exact restatements.  Floating-point
  reductions are accumulated in float64 and rounded once (the reference rounds
  in float32 pairwise order; the two agree to ~1 ulp of the block RMS and the
  tests state the tolerance).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# a1/a2  Hilbert coordinates            hilbert_quantization/core/hilbert_mapper.py:42-113
# ----------------------------------------------------------------------------

def is_pow2(n: int) -> bool:
    """core/hilbert_mapper.py:30 -- ``n > 0 and n & (n-1) == 0`` (n == 1 is accepted)."""
    return n > 0 and (n & (n - 1)) == 0


def d2xy(n: int, d) -> Tuple[np.ndarray, np.ndarray]:
    """Curve position -> (x, y); low-bits-first.  core/hilbert_mapper.py:42-66, rotate :92-113."""
    t = np.array(d, dtype=np.int64, copy=True).reshape(-1)
    x = np.zeros_like(t)
    y = np.zeros_like(t)
    s = 1
    while s < n:
        rx = 1 & (t >> 1)
        ry = 1 & (t ^ rx)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x = x + s * rx
        y = y + s * ry
        t >>= 2
        s <<= 1
    return x, y


def xy2d(n: int, x, y) -> np.ndarray:
    """(x, y) -> curve position; high-bits-first.  core/hilbert_mapper.py:68-90."""
    x = np.array(x, dtype=np.int64, copy=True).reshape(-1)
    y = np.array(y, dtype=np.int64, copy=True).reshape(-1)
    d = np.zeros_like(x)
    s = n // 2
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s //= 2
    return d


def hilbert_coordinates(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """All n*n coordinates in curve order.  core/hilbert_mapper.py:17-40."""
    if not is_pow2(n):
        raise ValueError(f"Grid size must be a power of 2, got {n}")
    return d2xy(n, np.arange(n * n, dtype=np.int64))


def hilbert_coordinates_list(n: int) -> List[Tuple[int, int]]:
    x, y = hilbert_coordinates(n)
    return list(zip(x.tolist(), y.tolist()))


_PERM_CACHE = {}


def stream_to_cell(n: int) -> np.ndarray:
    """cell[d] = y*n + x  (row-major cell of curve position d)."""
    p = _PERM_CACHE.get(n)
    if p is None:
        x, y = hilbert_coordinates(n)
        p = (y * n + x).astype(np.int64)
        _PERM_CACHE[n] = p
    return p


# ----------------------------------------------------------------------------
# a3/a4  map_to_2d / map_from_2d        core/hilbert_mapper.py:115-205
# ----------------------------------------------------------------------------

def map_to_2d(p: np.ndarray, dims: Tuple[int, int]) -> np.ndarray:
    """zeros((h, w), p.dtype); first len(p) curve cells filled.  core/hilbert_mapper.py:115-174."""
    w, h = dims
    if w != h:
        raise ValueError(f"Hilbert curve requires square dimensions, got {w}x{h}")
    if not is_pow2(w):
        raise ValueError(f"Dimension must be a power of 2, got {w}")
    if len(p) > w * h:
        raise ValueError(f"Too many parameters ({len(p)}) for dimensions {w}x{h} ({w * h} cells)")
    out = np.zeros(w * h, dtype=p.dtype)
    out[stream_to_cell(w)[: len(p)]] = p
    return out.reshape(h, w)


def map_to_2d_batch(p: np.ndarray, n: int) -> np.ndarray:
    """[N, D] -> [N, n, n]; same semantics row by row."""
    N, D = p.shape
    if D > n * n:
        raise ValueError("Too many parameters")
    out = np.zeros((N, n * n), dtype=p.dtype)
    out[:, stream_to_cell(n)[:D]] = p
    return out.reshape(N, n, n)


def map_from_2d(img: np.ndarray) -> np.ndarray:
    """All n*n values in curve order, dtype preserved, no trimming.  core/hilbert_mapper.py:176-205."""
    h, w = img.shape
    if w != h:
        raise ValueError(f"Hilbert curve requires square dimensions, got {w}x{h}")
    if not is_pow2(w):
        raise ValueError(f"Dimension must be a power of 2, got {w}")
    return img.reshape(-1)[stream_to_cell(w)].copy()


def map_from_2d_batch(imgs: np.ndarray) -> np.ndarray:
    N, h, w = imgs.shape
    return imgs.reshape(N, -1)[:, stream_to_cell(w)].copy()


# ----------------------------------------------------------------------------
# a5  dimensions                         core/dimension_calculator.py:36-128, config.py:18,27
# ----------------------------------------------------------------------------

VALID_DIMENSIONS = [4, 16, 64, 256, 1024, 4096, 16384]   # config.py:18 (cell counts)
MIN_EFFICIENCY_RATIO = 0.5                                # config.py:27


def nearest_power_of_4(value: int) -> int:
    """core/dimension_calculator.py:105-128."""
    if value <= 0:
        return 4
    for size in VALID_DIMENSIONS:
        if size >= value:
            return size
    power = VALID_DIMENSIONS[-1]
    while power < value:
        power *= 4
    return power


def optimal_dimensions(param_count: int) -> Tuple[int, int]:
    """core/dimension_calculator.py:36-61."""
    if param_count <= 0:
        raise ValueError("Parameter count must be positive")
    n = int(math.sqrt(nearest_power_of_4(param_count)))
    return (n, n)


def efficiency_ok(param_count: int, dims: Tuple[int, int], min_ratio: float = MIN_EFFICIENCY_RATIO) -> bool:
    """core/dimension_calculator.py:86-92 (the gate that rejects 1536 -> 64x64)."""
    return param_count / (dims[0] * dims[1]) >= min_ratio


def rag_optimal_dimensions(embedding_size: int) -> Tuple[int, int]:
    """rag/embedding_generation/generator.py:293-312 (no efficiency gate)."""
    side = math.ceil(math.sqrt(embedding_size))
    p = 1
    while p < side:
        p *= 2
    while p * p < embedding_size:
        p *= 2
    return (p, p)


# ----------------------------------------------------------------------------
# run-mean pyramid helpers (every index variant is run-means + a gather)
# ----------------------------------------------------------------------------

def run_means(stream: np.ndarray, k: int) -> np.ndarray:
    """R_k[j] = mean(stream[j*4^k:(j+1)*4^k]) accumulated in float64 (single rounding)."""
    run = 4 ** k
    m = stream.shape[-1] // run
    s = stream[..., : m * run].astype(np.float64).reshape(*stream.shape[:-1], m, run)
    return s.sum(axis=-1) / run


# ----------------------------------------------------------------------------
# a6  index variant B (streaming)        core/streaming_index_builder.py:45-102,154-243,315-343
# ----------------------------------------------------------------------------

B_MAX_LEVELS = 10   # streaming_index_builder.py:21


def b_levels_from_values(values: np.ndarray) -> List[np.ndarray]:
    """Per-level value lists after feeding ``values`` one by one.

    streaming_index_builder.py:70-102: level k+1 receives
    ``(w0 + w1 + w2 + w3) * 0.25`` (left-associated, Python float64) each time
    level k completes a group of four; incomplete trailing groups never promote.
    """
    cur = np.asarray(values, dtype=np.float64).reshape(-1)
    levels = [cur]
    for _ in range(1, B_MAX_LEVELS):
        m = cur.shape[0] // 4
        if m == 0:
            break
        g = cur[: 4 * m].reshape(m, 4)
        cur = (((g[:, 0] + g[:, 1]) + g[:, 2]) + g[:, 3]) * 0.25
        levels.append(cur)
    return levels


def b_level_allocations(level_sizes: Sequence[int], total_space: int) -> List[int]:
    """streaming_index_builder.py:207-243."""
    alloc = [0] * B_MAX_LEVELS
    non_empty = [k for k, sz in enumerate(level_sizes) if sz > 0]
    remaining = total_space
    for i, k in enumerate(non_empty):
        if i == len(non_empty) - 1:
            alloc[k] = remaining
        else:
            a = max(1, int(total_space * (0.5 ** (i + 1))))
            a = min(a, remaining)
            alloc[k] = a
            remaining -= a
    return alloc


def b_gather_plan(level_sizes: Sequence[int], total_space: int) -> List[Tuple[int, int]]:
    """(level, position) for each emitted value, in output order.  streaming_index_builder.py:177-194."""
    plan: List[Tuple[int, int]] = []
    alloc = b_level_allocations(level_sizes, total_space)
    for k, a in enumerate(alloc):
        if a <= 0 or k >= len(level_sizes):
            continue
        sz = level_sizes[k]
        if sz == 0:
            continue
        if sz > a:
            step = sz / a
            plan.extend((k, int(i * step)) for i in range(a))
        else:
            plan.extend((k, j) for j in range(sz))
    return plan


def index_b_from_values(values: np.ndarray, index_space_size: int) -> np.ndarray:
    """StreamingIndexBuilder fed with ``values`` then get_hierarchical_indices.  :154-205."""
    if index_space_size <= 0:
        return np.array([])
    levels = b_levels_from_values(values)
    sizes = [len(l) for l in levels] + [0] * (B_MAX_LEVELS - len(levels))
    plan = b_gather_plan(sizes, index_space_size)
    vals = np.array([levels[k][j] for k, j in plan], dtype=np.float64)
    out = np.zeros(index_space_size, dtype=np.float64)
    m = min(len(vals), index_space_size)
    out[:m] = vals[:m]
    return out


def index_b(image: np.ndarray, index_space_size: int) -> np.ndarray:
    """StreamingHilbertIndexGenerator.generate_optimized_indices.  :315-343 (float64 out)."""
    h, w = image.shape
    if w != h or not is_pow2(w):
        raise ValueError(f"Image must be square with power-of-2 dimensions, got {w}x{h}")
    return index_b_from_values(map_from_2d(image), index_space_size)


# ----------------------------------------------------------------------------
# a7  index variant A (traditional)      core/index_generator.py:34-219,313-356
# ----------------------------------------------------------------------------

def a_level_allocation(total_space: int) -> List[Tuple[int, int]]:
    """core/index_generator.py:34-98."""
    if total_space <= 0:
        return []
    allocations = []
    remaining = total_space
    max_grid = min(32, int(math.sqrt(total_space)))
    g = 1
    while g <= max_grid:
        g *= 2
    g //= 2
    g = max(g, 2)
    fraction = 0.5
    while remaining > 0 and g >= 1:
        need = g * g
        take = min(int(remaining * fraction), need, remaining)
        if take > 0:
            allocations.append((g, take))
            remaining -= take
        g //= 2
        fraction *= 0.5
        if fraction < 0.01:
            break
    if remaining > 0 and allocations:
        allocations.append((allocations[0][0], remaining))
    return allocations


def a_spatial_averages(image: np.ndarray, grid_size: int) -> np.ndarray:
    """Row-major block means.  core/index_generator.py:100-144."""
    if image.size == 0 or grid_size <= 0:
        return np.zeros(0, dtype=np.float64)
    h, w = image.shape
    sh, sw = h // grid_size, w // grid_size
    if sh == 0 or sw == 0:
        return np.array([image.astype(np.float64).mean()])
    blk = image[: sh * grid_size, : sw * grid_size].astype(np.float64)
    return blk.reshape(grid_size, sh, grid_size, sw).mean(axis=(1, 3)).reshape(-1)


def a_offset_sample_positions(h: int, w: int, section_size: int, available: int) -> List[Tuple[int, int]]:
    """(row, col) of every offset sample, in output order.  core/index_generator.py:146-219."""
    if h * w == 0 or section_size <= 0 or available <= 0:
        return []
    sy, sx = h // section_size, w // section_size
    if sy == 0 or sx == 0:
        pos = [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, w // 2)]
        return pos[:available]
    to_sample = min(available // 5, sy * sx)
    pos: List[Tuple[int, int]] = []
    count = 0
    for row in range(sy):
        for col in range(sx):
            if count >= to_sample:
                break
            r0, r1 = row * section_size, min((row + 1) * section_size, h)
            c0, c1 = col * section_size, min((col + 1) * section_size, w)
            pos += [(r0, c0), (r0, c1 - 1), (r1 - 1, c0), (r1 - 1, c1 - 1), ((r0 + r1) // 2, (c0 + c1) // 2)]
            count += 1
            if len(pos) >= available:
                break
        if len(pos) >= available:
            break
    return pos[:available]


def a_segments(h: int, w: int, index_space_size: int):
    """The per-allocation routing of _generate_traditional_indices (:313-356).

    Yields ("mean", grid, take) or ("sample", positions).  Mirrors the
    ``is_offset_sampling`` test at :329-332 literally (every entry after the
    first emitted one whose grid also appears in allocations[:-1]).
    """
    allocations = a_level_allocation(index_space_size)
    segs = []
    emitted = 0
    for grid, space in allocations:
        if space <= 0:
            continue
        is_offset = emitted > 0 and any(pg == grid for pg, _ in allocations[:-1])
        if is_offset:
            section = max(1, h // grid)
            pos = a_offset_sample_positions(h, w, section, space)
            segs.append(("sample", pos))
            emitted += len(pos)
        else:
            sh, sw = h // grid, w // grid
            count = 1 if (sh == 0 or sw == 0) else grid * grid
            take = min(count, space)
            segs.append(("mean", grid, take))
            emitted += take
    return segs


def index_a(image: np.ndarray, index_space_size: int) -> np.ndarray:
    """HierarchicalIndexGeneratorImpl._generate_traditional_indices -> float32[S].  :313-356."""
    if image.size == 0 or index_space_size <= 0:
        return np.array([])
    h, w = image.shape
    vals: List[float] = []
    for seg in a_segments(h, w, index_space_size):
        if seg[0] == "mean":
            _, grid, take = seg
            vals.extend(a_spatial_averages(image, grid)[:take].tolist())
        else:
            vals.extend(float(image[r, c]) for r, c in seg[1])
    res = np.array(vals[:index_space_size], dtype=np.float32)
    if len(res) < index_space_size:
        out = np.zeros(index_space_size, dtype=np.float32)
        out[: len(res)] = res
        res = out
    return res


def embed_index_row(image: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """core/index_generator.py:221-253."""
    if image.size == 0:
        return image
    h, w = image.shape
    out = np.zeros((h + 1, w), dtype=image.dtype)
    out[:h] = image
    m = min(len(indices), w)
    out[h, :m] = indices[:m]
    return out


def extract_index_row(enhanced: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """core/index_generator.py:255-290 (strips trailing zeros; keeps >= 1 element)."""
    if enhanced.size == 0:
        return enhanced, np.array([])
    h, w = enhanced.shape
    if h < 2:
        return enhanced, np.array([])
    row = enhanced[-1]
    nz = np.nonzero(row)[0]
    row = row[: nz[-1] + 1] if len(nz) else row[:1]
    return enhanced[:-1], row


# ----------------------------------------------------------------------------
# a8  index variant C (RAG multi-level)  rag/embedding_generation/hierarchical_index_generator.py
# ----------------------------------------------------------------------------

def nearest_pow2_le(n: int) -> int:
    """hierarchical_index_generator.py:557-567."""
    if n <= 0:
        return 1
    p = 1
    while p * 2 <= n:
        p *= 2
    return p


def c_granularity_levels(width: int, min_granularity: int = 2, max_index_rows: int = 8) -> List[int]:
    """hierarchical_index_generator.py:23-68."""
    finest = max(min_granularity, int(math.sqrt(width)))
    finest = nearest_pow2_le(finest)
    levels = []
    g = finest
    while g >= min_granularity and len(levels) < max_index_rows:
        levels.append(g)
        g //= 2
    return levels


def c_section_order(g: int) -> List[Tuple[int, int]]:
    """(row, col) visiting order of _generate_hilbert_coordinates(g).  :286-342.

    g == 1 -> [(0,0)]; g == 2 is hard-coded [(0,0),(0,1),(1,1),(1,0)] as
    (row, col); for g >= 4 the reference's recursive generator equals the core
    d2xy(g) sequence read as (row, col) = (y, x) (pinned for g <= 64 in
    oracle/pin_against_reference.py).
    """
    if g == 1:
        return [(0, 0)]
    if g == 2:
        return [(0, 0), (0, 1), (1, 1), (1, 0)]
    if g & (g - 1):
        g = nearest_pow2_le(g)
    x, y = hilbert_coordinates(g)
    return list(zip(y.tolist(), x.tolist()))


def c_index_row(image: np.ndarray, g: int) -> np.ndarray:
    """_calculate_hilbert_order_averages.  :204-244."""
    h, w = image.shape
    sh, sw = h // g, w // g
    if sh == 0 or sw == 0:
        return np.array([image.astype(np.float64).mean()]).astype(image.dtype)
    order = c_section_order(g)
    gg = nearest_pow2_le(g) if (g & (g - 1)) else g
    blk = image[: sh * gg, : sw * gg].astype(np.float64).reshape(gg, sh, gg, sw).mean(axis=(1, 3))
    rows = np.array([r for r, _ in order])
    cols = np.array([c for _, c in order])
    return blk[rows, cols].astype(image.dtype)


def index_c_rows(image: np.ndarray) -> List[np.ndarray]:
    """create_progressive_granularity_levels.  :148-178."""
    h, w = image.shape
    return [c_index_row(image, g) for g in c_granularity_levels(w)]


def index_c(image: np.ndarray) -> np.ndarray:
    """generate_multi_level_indices: (H, W) -> (H + L, W), same dtype.  :103-146."""
    if image.ndim != 2:
        raise ValueError("Embedding image must be 2D")
    h, w = image.shape
    rows = index_c_rows(image)
    out = np.zeros((h + len(rows), w), dtype=image.dtype)
    out[:h] = image
    for i, r in enumerate(rows):
        if len(r) <= w:
            out[h + i, : len(r)] = r
        else:
            out[h + i, :] = r[:w]
    return out


def index_c_batch_compact(grids: np.ndarray) -> np.ndarray:
    """[N, n, n] (square pow2) -> [N, sum(g^2)] compact rows, levels concatenated finest first."""
    N, n, _ = grids.shape
    streams = map_from_2d_batch(grids)
    parts = []
    for g in c_granularity_levels(n):
        k = int(round(math.log2(n // g)))
        r = run_means(streams, k)                      # [N, g*g] in curve order
        if g == 2:
            r = r[:, [0, 3, 2, 1]]
        parts.append(r.astype(grids.dtype))
    return np.concatenate(parts, axis=1)


def c_extract_rows(enhanced: np.ndarray, original_height: int) -> Tuple[np.ndarray, List[np.ndarray]]:
    """extract_indices_from_image with an explicit original_height.  :387-441."""
    height = enhanced.shape[0]
    oh = max(0, min(original_height, height))
    if oh >= height:
        return enhanced, []
    rows = []
    for r in range(oh, height):
        row = enhanced[r]
        nz = np.nonzero(row)[0]
        rows.append(row[: nz[-1] + 1] if len(nz) else row[:1])
    return enhanced[:oh], rows


# ----------------------------------------------------------------------------
# a10  uint8 quantise / dequantise        core/compressor.py:256-303
# ----------------------------------------------------------------------------

def normalize_u8(image: np.ndarray):
    """_normalize_for_compression: truncating cast; constant image -> 128.  Returns (u8, min, max)."""
    mn, mx = image.min(), image.max()
    if mx == mn:
        return np.full_like(image, 128, dtype=np.uint8), mn, mx
    return ((image - mn) / (mx - mn) * 255).astype(np.uint8), mn, mx


def denormalize_u8(q: np.ndarray, mn, mx) -> np.ndarray:
    """_denormalize_from_compression.  compressor.py:282-303."""
    if mx == mn:
        return np.full_like(q, mn, dtype=np.float32)
    return (q.astype(np.float32) / 255.0) * (mx - mn) + mn


# ----------------------------------------------------------------------------
# a12  RAG progressive filter              rag/search/engine.py:51-95,178-287,1025-1051
# ----------------------------------------------------------------------------

def stripped_length(rows: np.ndarray) -> np.ndarray:
    """Length after trailing-zero strip, minimum 1.  hierarchical_index_generator.py:424-439."""
    nz = rows != 0
    W = rows.shape[-1]
    last = W - np.argmax(nz[..., ::-1], axis=-1)
    return np.where(nz.any(axis=-1), last, 1).astype(np.int64)


def rag_threshold(level: int) -> float:
    """rag/search/engine.py:262-268."""
    return min(0.3 + 0.1 * (3 - min(level, 3)), 0.8)


def rag_ratio(level: int) -> float:
    """rag/search/engine.py:272-277."""
    return 0.3 if level == 0 else 0.5 if level == 1 else 0.7


def rag_level_scores(q_row: np.ndarray, q_len: int, c_rows: np.ndarray, c_len: np.ndarray) -> np.ndarray:
    """(cos + 1)/2 on the common prefix of stripped rows; 0 when a norm is 0.

    rag/search/engine.py:205-227 + :1025-1051.  ``q_row``/``c_rows`` are full
    zero-padded rows ([W], [N, W]); ``q_len``/``c_len`` the stripped lengths.
    """
    W = c_rows.shape[1]
    m = np.minimum(q_len, c_len)                                   # [N]
    pos = np.arange(W)[None, :]
    inside = pos < m[:, None]
    q = q_row.astype(np.float64)[None, :] * inside
    c = c_rows.astype(np.float64) * inside
    dot = (q * c).sum(1)
    nq = np.sqrt((q * q).sum(1))
    nc = np.sqrt((c * c).sum(1))
    ok = (nq > 0) & (nc > 0)
    cos = np.where(ok, dot / np.where(ok, nq * nc, 1.0), 0.0)
    return np.where(ok, (cos + 1.0) / 2.0, 0.0)


def rag_progressive_filter(q_rows: Sequence[np.ndarray], c_rows: Sequence[np.ndarray],
                           return_scores: bool = False):
    """progressive_hierarchical_search with rows fed explicitly.

    ``q_rows[l]`` is the zero-padded query row of level l ([W]); ``c_rows[l]``
    the candidates' rows ([N, W]).  Levels are the rows in order (row 0 =
    finest).  Returns the survivor ids in the reference's final order (sorted
    by last-level score, stable).  rag/search/engine.py:80-95,178-287.
    """
    N = c_rows[0].shape[0]
    cand = np.arange(N, dtype=np.int64)
    trace = []
    for level, (q, c) in enumerate(zip(q_rows, c_rows)):
        if cand.size == 0:
            break
        q_len = int(stripped_length(q[None, :])[0])
        c_sub = c[cand]
        s = rag_level_scores(q, q_len, c_sub, stripped_length(c_sub))
        order = np.argsort(-s, kind="stable")
        cap = max(1, int(len(cand) * rag_ratio(level)))
        thr = rag_threshold(level)
        keep = order[:cap]
        keep = keep[s[keep] >= thr]
        trace.append((cand.copy(), s.copy(), thr, cap))
        cand = cand[keep]
    if return_scores:
        return cand, trace
    return cand


# ----------------------------------------------------------------------------
# a13/a15  cosine rerank + top-k            rag/search/engine.py:622-660, :512
# ----------------------------------------------------------------------------

def cosine01(q: np.ndarray, c: np.ndarray) -> np.ndarray:
    """(dot/(|q||c|) + 1)/2, 0 when a norm is 0; q [D], c [N, D] (float64 accumulate)."""
    q64 = q.astype(np.float64)
    c64 = c.astype(np.float64)
    dot = c64 @ q64
    nq = math.sqrt(float(q64 @ q64))
    nc = np.sqrt(np.einsum("nd,nd->n", c64, c64))
    ok = (nc > 0) & (nq > 0)
    cos = np.where(ok, dot / np.where(ok, nq * nc, 1.0), 0.0)
    return np.where(ok, (cos + 1.0) / 2.0, 0.0)


def topk_stable(ids: np.ndarray, scores: np.ndarray, k: int):
    """Stable descending sort, first k (ties -> earlier list position).  rag/search/engine.py:512."""
    order = np.argsort(-scores, kind="stable")[:k]
    return ids[order], scores[order]


def progressive_search(query: np.ndarray, db: np.ndarray, n: int, k: int,
                       db_rows: Optional[List[np.ndarray]] = None):
    """North-star progressive search of one query embedding against db [N, D].

    Composition of SURVEY 8c: map_to_2d -> index C rows -> progressive filter
    -> cosine01 rerank of survivors -> stable top-k.  Survivors are re-ranked in
    ascending-id order so that exact ties resolve to the lower id.
    """
    W = n
    levels = c_granularity_levels(n)
    if db_rows is None:
        compact = index_c_batch_compact(map_to_2d_batch(db, n))
        db_rows, o = [], 0
        for g in levels:
            r = np.zeros((db.shape[0], W), dtype=db.dtype)
            r[:, : g * g] = compact[:, o:o + g * g]
            db_rows.append(r)
            o += g * g
    qc = index_c_batch_compact(map_to_2d_batch(query[None, :], n))[0]
    q_rows, o = [], 0
    for g in levels:
        r = np.zeros(W, dtype=query.dtype)
        r[: g * g] = qc[o:o + g * g]
        q_rows.append(r)
        o += g * g
    surv = np.sort(rag_progressive_filter(q_rows, db_rows))
    if surv.size == 0:
        return surv, np.zeros(0)
    s = cosine01(query, db[surv])
    return topk_stable(surv, s, k)


# ----------------------------------------------------------------------------
# a14  comprehensive similarity blend        rag/search/engine.py:516-575, :662-727, :1053-1138
# ----------------------------------------------------------------------------

def granularity_weights(num_levels: int) -> np.ndarray:
    """8^(L-i-1), normalised, first level doubled, renormalised.  rag/search/engine.py:1101-1138."""
    if num_levels <= 0:
        return np.array([])
    if num_levels == 1:
        return np.array([1.0])
    w = np.array([8.0 ** (num_levels - i - 1) for i in range(num_levels)])
    w = w / w.sum()
    w[0] *= 2.0
    return w / w.sum()


def _cos01_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(cos + 1)/2 along the last axis, 0 where a norm is 0 (float64 accumulate)."""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    dot = (a64 * b64).sum(-1)
    na, nb = np.sqrt((a64 * a64).sum(-1)), np.sqrt((b64 * b64).sum(-1))
    ok = (na > 0) & (nb > 0)
    return np.where(ok, (dot / np.where(ok, na * nb, 1.0) + 1.0) / 2.0, 0.0)


def spatial_locality_similarity(q_grid: np.ndarray, c_grids: np.ndarray) -> np.ndarray:
    """Mean of (cos + 1)/2 over ws x ws windows at stride ws // 2, ws = min(4, H // 4, W // 4); plain
    cosine of the grids when ws < 2.  q_grid [H, W], c_grids [N, H, W].  rag/search/engine.py:662-714."""
    N, H, W = c_grids.shape
    ws = min(4, H // 4, W // 4)
    if ws < 2:
        return _cos01_rows(np.broadcast_to(q_grid.reshape(1, -1), (N, H * W)), c_grids.reshape(N, -1))
    step = ws // 2
    acc = np.zeros(N)
    count = 0
    for i in range(0, H - ws + 1, step):
        for j in range(0, W - ws + 1, step):
            qw = q_grid[i:i + ws, j:j + ws].reshape(1, -1)
            cw = c_grids[:, i:i + ws, j:j + ws].reshape(N, -1)
            acc += _cos01_rows(np.broadcast_to(qw, cw.shape), cw)
            count += 1
    return acc / count if count else np.zeros(N)


def comprehensive_similarity(q_frame: np.ndarray, c_frames: np.ndarray, original_height: int) -> np.ndarray:
    """0.5 * hierarchical + 0.3 * embedding cosine + 0.2 * spatial locality for one query frame against N
    candidate frames (enhanced frames: rows [0, H) = grid, rows [H, H + L) = index rows, read explicitly as
    in SURVEY 8c).  Index rows that are all zero in the query or a candidate are NOT dropped here (the
    reference's own extractor drops them and thereby shifts the later levels; dense data never hits that).
    rag/search/engine.py:516-575 (+ :994-1099 for the hierarchical part, weights :1101-1138)."""
    H = original_height
    N = c_frames.shape[0]
    L = q_frame.shape[0] - H
    if L > 0:
        w = granularity_weights(L)
        hier = np.zeros(N)
        for l in range(L):
            hier += w[l] * _cos01_rows(np.broadcast_to(q_frame[H + l][None, :], (N, q_frame.shape[1])), c_frames[:, H + l, :])
        hier /= w.sum()
    else:
        hier = np.zeros(N)
    emb = _cos01_rows(np.broadcast_to(q_frame[:H].reshape(1, -1), (N, H * q_frame.shape[1])), c_frames[:, :H].reshape(N, -1))
    spatial = spatial_locality_similarity(q_frame[:H], c_frames[:, :H])
    return 0.5 * hier + 0.3 * emb + 0.2 * spatial


# ----------------------------------------------------------------------------
# f1  PrecomputedHilbertIndexer               core/precomputed_hilbert_index.py:122-212
# ----------------------------------------------------------------------------

def precomputed_granularity_levels(image_size: int, max_levels: int = 6, min_square_size: int = 2) -> List[Tuple[int, int]]:
    """(grid_size, square_size) per level.  core/precomputed_hilbert_index.py:122-150."""
    levels = []
    square = min_square_size
    while square <= image_size // 2 and len(levels) < max_levels:
        grid = image_size // square
        if grid >= 2:
            levels.append((grid, square))
        square *= 2
    if len(levels) == 0 or levels[-1][1] < image_size:
        levels.append((1, image_size))
    return levels


def precomputed_level_averages(image: np.ndarray, grid_size: int, square_size: int) -> np.ndarray:
    """Aligned square means (row-major) followed by the squares offset by half a side (row-major), float32.
    core/precomputed_hilbert_index.py:152-212."""
    h, w = image.shape
    out = []
    for r in range(grid_size):
        for c in range(grid_size):
            y, x = r * square_size, c * square_size
            reg = image[y:min(y + square_size, h), x:min(x + square_size, w)]
            if reg.size:
                out.append(float(np.mean(reg)))
    off = square_size // 2
    if off > 0:
        for r in range(grid_size - 1):
            for c in range(grid_size - 1):
                y, x = r * square_size + off, c * square_size + off
                reg = image[y:min(y + square_size, h), x:min(x + square_size, w)]
                if reg.size:
                    out.append(float(np.mean(reg)))
    return np.array(out, dtype=np.float32)


# ----------------------------------------------------------------------------
# f4  video-path hierarchical similarity / frame ordering     core/video_storage.py:741-781, :1203-1277, :1751-1803
# ----------------------------------------------------------------------------

def video_hierarchical_similarity(q: np.ndarray, c: np.ndarray) -> float:
    """(pearson + 1) / 2 on the common prefix, clamped; zero-variance special case.  core/video_storage.py:763-781."""
    if len(q) == 0 or len(c) == 0:
        return 0.0
    m = min(len(q), len(c))
    a, b = q[:m], c[:m]
    if np.std(a) == 0 or np.std(b) == 0:
        return 1.0 if np.allclose(a, b) else 0.0
    corr = np.corrcoef(a, b)[0, 1]
    return max(0.0, min(1.0, (corr + 1.0) / 2.0))


def video_sort_frames(frame_indices: Sequence[np.ndarray]) -> List[int]:
    """Greedy nearest-neighbour order starting at the frame closest to the centroid.  core/video_storage.py:1203-1277."""
    F = len(frame_indices)
    if F <= 1:
        return list(range(F))
    centroid = np.mean([f for f in frame_indices if len(f) > 0], axis=0)
    best, best_d = None, float("inf")
    for i, f in enumerate(frame_indices):
        d = np.linalg.norm(f - centroid)
        if d < best_d:
            best, best_d = i, d
    order = [best]
    remaining = [i for i in range(F) if i != best]
    while remaining:
        bj, bs = None, -1.0
        for j in remaining:
            s = video_hierarchical_similarity(frame_indices[order[-1]], frame_indices[j])
            if s > bs:
                bj, bs = j, s
        order.append(bj)
        remaining.remove(bj)
    return order


def video_insertion_position(new_indices: np.ndarray, existing: Sequence[np.ndarray]) -> int:
    """core/video_storage.py:1751-1803."""
    if not len(existing):
        return 0
    sims = [video_hierarchical_similarity(new_indices, e) for e in existing]
    pos, best = 0, -1.0
    if sims[0] > best:
        best, pos = sims[0], 0
    for i in range(len(sims) - 1):
        sc = (sims[i] + sims[i + 1]) / 2.0
        if sc > best:
            best, pos = sc, i + 1
    if sims[-1] > best:
        best, pos = sims[-1], len(sims)
    return pos


def video_hierarchical_search(q: np.ndarray, frames: Sequence, max_results: int, similarity_threshold: float = 0.1):
    """core/video_search.py:215-262 with :1316-1328: finest-level comparison of the core engine per frame, strict
    threshold, stable descending sort.  Returns [(frame position, similarity)]."""
    cands = []
    for i, f in enumerate(frames):
        if f is None:
            continue
        if len(q) == 0 or len(f) == 0:
            s = 0.0
        else:
            ql, cl = core_parse_levels(len(q), len(q)), core_parse_levels(len(f), len(f))
            if not ql or not cl:
                s = 0.0
            else:
                a, b = q[ql[0][1]:ql[0][2]], np.asarray(f)[cl[0][1]:cl[0][2]]
                m = min(len(a), len(b))
                s = float(core_level_similarity(np.asarray(a[:m], dtype=np.float64), np.asarray(b[:m], dtype=np.float64)[None, :])[0]) if m else 0.0
        if s > similarity_threshold:
            cands.append((i, s))
    cands.sort(key=lambda t: t[1], reverse=True)
    return cands[:max_results]


# ----------------------------------------------------------------------------
# a11  core progressive search               core/search_engine.py:42-388
# ----------------------------------------------------------------------------

def core_parse_levels(length: int, total_space: int) -> List[Tuple[int, int, int, bool]]:
    """_parse_index_structure -> [(grid, start, end, is_offset)].  core/search_engine.py:42-109."""
    if length == 0 or total_space <= 0:
        return []
    levels = []
    remaining = total_space
    cur = 0
    max_grid = min(32, int(math.sqrt(total_space)))
    g = 1
    while g <= max_grid:
        g *= 2
    g //= 2
    g = max(g, 2)
    fraction = 0.5
    seen = set()
    while remaining > 0 and g >= 1 and cur < length:
        need = g * g
        take = min(int(remaining * fraction), need, remaining)
        if take > 0:
            levels.append((g, cur, cur + take, g in seen))
            seen.add(g)
            cur += take
            remaining -= take
        g //= 2
        fraction *= 0.5
        if fraction < 0.01:
            break
    if remaining > 0 and cur < length and levels:
        levels.append((levels[0][0], cur, min(cur + remaining, length), True))
    return levels


def core_level_similarity(q: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Per-level similarity of one query slice q [m] against candidate slices c [N, m].

    core/search_engine.py:151-189: population std; both constant -> 1.0 iff the
    means differ by < 1e-6 else 0.0; one constant -> 0.1; else
    clamp(0.7*(corr+1)/2 + 0.3*max(0, 1 - mse/(E[q^2]+E[c^2]))).
    """
    q = q.astype(np.float64)
    c = c.astype(np.float64)
    qs, qm = q.std(), q.mean()
    cs, cm = c.std(axis=1), c.mean(axis=1)
    both_const = (qs == 0) & (cs == 0)
    one_const = ((qs == 0) | (cs == 0)) & ~both_const
    safe_cs = np.where(cs == 0, 1.0, cs)
    safe_qs = qs if qs != 0 else 1.0
    qn = (q - qm) / safe_qs
    cn = (c - cm[:, None]) / safe_cs[:, None]
    corr = (qn[None, :] * cn).mean(axis=1)
    sim = (corr + 1.0) / 2.0
    mse = ((q[None, :] - c) ** 2).mean(axis=1)
    mx = (q ** 2).mean() + (c ** 2).mean(axis=1)
    dist = np.where(mx > 0, np.maximum(0.0, 1.0 - mse / np.where(mx > 0, mx, 1.0)), 1.0)
    out = np.clip(0.7 * sim + 0.3 * dist, 0.0, 1.0)
    out = np.where(one_const, 0.1, out)
    out = np.where(both_const, np.where(np.abs(qm - cm) < 1e-6, 1.0, 0.0), out)
    return out


def core_all_level_similarities(q: np.ndarray, c: np.ndarray) -> np.ndarray:
    """[N, n_levels] per-level similarities; query and candidates share one length S."""
    S = q.shape[0]
    levels = core_parse_levels(S, S)
    out = np.zeros((c.shape[0], len(levels)))
    for i, (_, a, b, _) in enumerate(levels):
        out[:, i] = core_level_similarity(q[a:b], c[:, a:b])
    return out


def core_progressive_search(q: np.ndarray, c: np.ndarray, max_results: int,
                            similarity_threshold: float = 0.1, max_candidates_per_level: int = 100):
    """progressive_search over candidates with equal-length index vectors.

    core/search_engine.py:232-300 (filter) + :340-388 (final weighted score).
    Returns (ids, scores) in the reference's result order.
    """
    N = c.shape[0]
    if q.shape[0] == 0 or N == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    sims = core_all_level_similarities(q, c)
    L = sims.shape[1]
    if L == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    w = 1.0 / (np.arange(L) + 1.0)
    cur = np.arange(N, dtype=np.int64)
    for lvl in range(L):
        if len(cur) <= max_candidates_per_level:
            break
        ls = sims[cur, lvl]
        combined = (sims[cur, : lvl + 1] * w[: lvl + 1]).sum(1) / w[: lvl + 1].sum()
        keep = ls >= similarity_threshold
        kept, kc = cur[keep], combined[keep]
        order = np.argsort(-kc, kind="stable")[:max_candidates_per_level]
        nxt = kept[order]
        if len(nxt) == 0 and len(cur) > 0:
            nxt = cur[[int(np.argmax(ls))]]
        cur = nxt
    overall = np.clip((sims[cur] * w).sum(1) / w.sum(), 0.0, 1.0)
    order = np.argsort(-overall, kind="stable")[:max_results]
    return cur[order], overall[order]
