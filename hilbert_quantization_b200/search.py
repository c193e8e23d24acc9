"""Progressive search on the device: coarse index filter -> cosine rerank -> top-k.

Reference semantics: rag/search/engine.py:51-95, :178-287 (filter), :622-660 (cosine),
:512 (stable sort) with index rows fed explicitly, and
core/search_engine.py:23-388 for the core engine over QuantizedModel indices.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _device as dev
from . import plans
from ._lib import IndexLayout, check, lib
from .dimension import rag_optimal_dimensions
from .index import map_and_index


class PhaseTimer:
    """Optional CUDA-event timing of the phases of search_batch (bench.py installs one in
    `PHASE_TIMER`; events are recorded on the stream the kernels are launched on)."""

    def __init__(self):
        self.events = []

    def start(self, name):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        self.events.append((name, e0, e1))
        return e1

    @staticmethod
    def stop(e1):
        e1.record(torch.cuda.current_stream())

    def totals_ms(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.events:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out


PHASE_TIMER: Optional[PhaseTimer] = None


def _phase(name):
    return PHASE_TIMER.start(name) if PHASE_TIMER is not None else None


def _end(tok):
    if tok is not None:
        PhaseTimer.stop(tok)


def rag_threshold(level: int) -> float:
    """rag/search/engine.py:262-268"""
    base_threshold = 0.3
    level_factor = 0.1
    return min(base_threshold + (level_factor * (3 - min(level, 3))), 0.8)


def rag_ratio(level: int) -> float:
    """rag/search/engine.py:272-277"""
    return 0.3 if level == 0 else (0.5 if level == 1 else 0.7)


def xstar_for(thr: float) -> np.float32:
    """Smallest float32 x whose float32-evaluated score (x + 1) * 0.5 reaches `thr` (compared as
    float64, like `score >= threshold` in rag/search/engine.py:284): the cosine cut of a level.
    Bisection over the ordered float32 bit patterns (the score is monotone in x)."""
    one, half = np.float32(1.0), np.float32(0.5)

    def from_ord(i: int) -> np.float32:          # order-preserving int -> float32
        bits = i if i >= 0 else (-(i + 1)) | 0x80000000
        return np.array([bits & 0xffffffff], dtype=np.uint32).view(np.float32)[0]

    def ok(i: int) -> bool:
        return float((from_ord(i) + one) * half) >= thr
    lo, hi = -0x40400000, 0x40400000              # -3.0 .. 3.0
    if ok(lo):
        return from_ord(lo)
    if not ok(hi):
        return np.float32(np.inf)
    while hi - lo > 1:                            # invariant: not ok(lo), ok(hi)
        mid = (lo + hi) // 2
        if ok(mid):
            hi = mid
        else:
            lo = mid
    return np.float32(from_ord(hi))


def make_layout(n: int, D: int) -> Tuple[IndexLayout, List[int]]:
    """Compact variant-C layout of an n x n grid holding D real values."""
    levels = plans.c_levels(n)
    lay = IndexLayout()
    lay.L = len(levels)
    off = 0
    for i, g in enumerate(levels):
        w = min(g * g, n)
        run = (n // g) ** 2                       # curve positions per section
        lay.lvl_off[i] = off
        lay.lvl_w[i] = w
        keff = min(w, max(1, -(-D // run)))
        if g == 2:                                # the g = 2 row is stored as quadrants [0, 3, 2, 1]
            keff = w if D > run else 1
        lay.lvl_keff[i] = keff
        off += w
    lay.Lsum = off
    return lay, levels


class EmbeddingDatabase:
    """Device-resident shard of the embedding database.

    Holds, per row: the float32 embedding, its L2 norm, the compact variant-C index rows
    and their stripped lengths.  Built by ONE fused map+index pass over the embeddings
    (the 2-D grids are produced only if `keep_grids`).  `EmbeddingDatabase.from_chunks` builds a bf16-only shard
    (no float32 rows) from a stream of row blocks."""

    MAX_EXCEPTION_ROWS = 4096

    def __init__(self, embeddings, n: Optional[int] = None, device=None, keep_grids: bool = False, id_base: int = 0,
                 bf16: bool = True, tc_filter: bool = True):
        d = dev.require_cuda(device if device is not None else (embeddings.device if isinstance(embeddings, torch.Tensor)
                                                                 and embeddings.is_cuda else None))
        self.device = d
        self.emb = dev.f32_device(embeddings, d)
        if self.emb.dim() != 2:
            raise ValueError("embeddings must be [N, D]")
        self.N, self.D = self.emb.shape
        self.n = int(n) if n is not None else rag_optimal_dimensions(self.D)[0]
        self.id_base = int(id_base)
        self.layout, self.levels = make_layout(self.n, self.D)
        # operand of the tensor-core rerank: unit rows (the epilogue then needs no 1 / |c|); zero-norm rows score 0.0 and
        # are listed separately
        fused = None if keep_grids else shard_ingest(self.emb, self.n, want_bf16=bf16)
        if fused is not None:                       # one pass over the fp32 rows: index rows + norms + bf16 unit rows
            self.grids = None
            self.idx, self.norms, self.emb_bf16 = fused
        else:
            grids, self.idx = map_and_index(self.emb, self.n, variant="C", layout="compact", want_grid=keep_grids)
            self.grids = grids if keep_grids else None
            self.norms = row_norms(self.emb)
            self.emb_bf16 = to_bf16(self.emb, self.norms) if bf16 else None
        # database-side term of the rerank guard's error bound: max_r |bf16 unit row - c_r / |c_r||_2  (hq_rerank_topk_unit_bf16)
        self.dc_max = 0.0
        if self.emb_bf16 is not None and self.N > 0:
            worst = torch.zeros(1, dtype=torch.float32, device=d)
            with torch.cuda.device(d):
                check(lib.hq_bf16_unit_error_max(dev.ptr(self.emb), self.N, self.D, self.emb.stride(0), dev.ptr(self.norms),
                                                 dev.ptr(self.emb_bf16), self.emb_bf16.stride(0), dev.ptr(worst), dev.stream_ptr()))
            self.dc_max = float(worst.item())
        self._finish_build(bf16, tc_filter)

    @classmethod
    def from_chunks(cls, chunks, N: int, D: int, n: Optional[int] = None, device=None, id_base: int = 0,
                    tc_filter: bool = True) -> "EmbeddingDatabase":
        """bf16-ONLY shard built from a stream of float32 row blocks (device tensors [m, D], in row order, N rows in
        total): every block goes through hq_shard_ingest once and is dropped; the shard keeps the bf16 unit rows, the
        index rows and the filter operands -- 2 D + ~600 bytes per row instead of 6 D (a 100 M x 768 database over two or
        more GPUs, BASELINE config 5).  The stored rows ARE the bf16 unit rows: scores are (cos(q, stored row) + 1) / 2,
        exact fp32 for the stored values (include/hq_b200.h, hq_rerank_topk_unit_bf16); the filter runs on the fp32 index
        rows of the original embeddings, exactly like the full database."""
        d = dev.require_cuda(device)
        self = cls.__new__(cls)
        self.device, self.N, self.D = d, int(N), int(D)
        self.n = int(n) if n is not None else rag_optimal_dimensions(self.D)[0]
        self.id_base = int(id_base)
        self.layout, self.levels = make_layout(self.n, self.D)
        self.emb = self.grids = None
        Lsum = int(self.layout.Lsum)
        pitch = (self.D + 7) // 8 * 8
        self.idx = torch.empty((self.N, Lsum), dtype=torch.float32, device=d)
        self.emb_bf16 = torch.empty((self.N, pitch), dtype=torch.bfloat16, device=d)
        if pitch != self.D:
            self.emb_bf16.zero_()
        at = 0
        for block in chunks:
            block = dev.f32_device(block, d)
            m = int(block.shape[0])
            if block.dim() != 2 or block.shape[1] != self.D or at + m > self.N:
                raise ValueError("chunks must be [m, D] row blocks adding up to N rows")
            fused = shard_ingest(block, self.n, want_bf16=True)
            if fused is not None:
                idx, _, unit = fused
            else:
                _, idx = map_and_index(block, self.n, variant="C", layout="compact", want_grid=False)
                unit = to_bf16(block, row_norms(block))
            self.idx[at:at + m] = idx
            self.emb_bf16[at:at + m, : unit.shape[1]] = unit
            at += m
            del block, fused, idx, unit
        if at != self.N:
            raise ValueError(f"chunks held {at} rows, expected {self.N}")
        # norms of the STORED rows (1 up to the bf16 rounding; 0 for zero rows) and the guard term that covers them
        self.norms = torch.empty(self.N, dtype=torch.float32, device=d)
        step = 1 << 20
        worst = 0.0
        for s0 in range(0, self.N, step):
            blk = self.emb_bf16[s0:s0 + step, : self.D].to(torch.float32)
            nb = row_norms(blk)
            self.norms[s0:s0 + step] = nb
            nz = nb[nb > 0]
            if nz.numel():
                worst = max(worst, float((nz - 1.0).abs().max().item()))
            del blk, nb, nz
        self.dc_max = worst * 1.01 + 1e-6
        self._finish_build(True, tc_filter)
        return self

    @classmethod
    def from_frames(cls, batch, device=None, id_base: int = 0, bf16: bool = True, tc_filter: bool = True) -> "EmbeddingDatabase":
        """The search shard of a batch of stored ENHANCED FRAMES (frames.EmbeddingFrameBatch: grid + one index row per
        level, the reference's storage format, rag/.../hierarchical_index_generator.py:344-385): the embeddings come back
        through map_from_2d, the index rows are the frames' own rows (sliced, not recomputed), the original height is the
        batch's explicit field (the reference guesses it from the pixel values, rag/search/engine.py:134-162)."""
        d = dev.require_cuda(device if device is not None else batch.frames.device)
        self = cls.__new__(cls)
        self.device = d
        self.emb = batch.embeddings().to(d)
        self.N, self.D = (int(x) for x in self.emb.shape)
        self.n = int(batch.original_height)
        self.id_base = int(id_base)
        self.layout, self.levels = make_layout(self.n, self.D)
        self.grids = None
        self.idx = batch.index_rows().to(d)
        if int(self.idx.shape[1]) != int(self.layout.Lsum):
            raise ValueError("frame index rows do not match the index layout of this grid size")
        self.norms = row_norms(self.emb)
        self.emb_bf16 = to_bf16(self.emb, self.norms) if bf16 else None
        self.dc_max = 0.0
        if self.emb_bf16 is not None and self.N > 0:
            worst = torch.zeros(1, dtype=torch.float32, device=d)
            with torch.cuda.device(d):
                check(lib.hq_bf16_unit_error_max(dev.ptr(self.emb), self.N, self.D, self.emb.stride(0), dev.ptr(self.norms),
                                                 dev.ptr(self.emb_bf16), self.emb_bf16.stride(0), dev.ptr(worst), dev.stream_ptr()))
            self.dc_max = float(worst.item())
        self._finish_build(bf16, tc_filter)
        return self

    def _finish_build(self, bf16: bool, tc_filter: bool):
        d = self.device
        self.lens = row_lengths(self.idx, self.layout)
        self.zero_rows = (self.norms == 0).nonzero().flatten().to(torch.int32) if bf16 else None
        # fast filter: per-level row norms + "every stored length is the structural one" check
        self.level_norms = torch.empty((self.N, int(self.layout.L)), dtype=torch.float32, device=d)
        flag = torch.zeros(1, dtype=torch.int32, device=d)
        with torch.cuda.device(d):
            check(lib.hq_filter_level_norms(dev.ptr(self.idx), dev.ptr(self.lens), self.N, C.byref(self.layout),
                                            dev.ptr(self.level_norms), dev.ptr(flag), dev.stream_ptr()))
        self._keff = torch.tensor([int(self.layout.lvl_keff[l]) for l in range(int(self.layout.L))], dtype=torch.int16, device=d)
        # Rows whose stored length differs from the structural one (a block mean that is exactly 0.0 at the end of
        # an index row: about one row in 10 M at 768-D) are "exceptions": masked out of the dense filter passes and
        # scored pair by pair.  Truly sparse data (many such rows) uses the exact path.
        self.exc_rows = torch.empty(0, dtype=torch.int32, device=d)
        if self.N > 0 and int(flag.item()) != 0:
            self.exc_rows = (self.lens != self._keff).any(dim=1).nonzero().flatten().to(torch.int32)
        self.fast_filter_ok = bool(lib.hq_filter_fast_supported(C.byref(self.layout))) and self.exc_rows.numel() <= self.MAX_EXCEPTION_ROWS
        self._xstar = torch.tensor([float(xstar_for(rag_threshold(l))) for l in range(int(self.layout.L))], dtype=torch.float32, device=d)
        self._ratio = torch.tensor([rag_ratio(l) for l in range(int(self.layout.L))], dtype=torch.float64, device=d)
        self._ratio_host = (C.c_double * 8)(*([rag_ratio(l) for l in range(int(self.layout.L))] + [1.0] * (8 - int(self.layout.L))))
        self._thr_host = (C.c_double * 3)(*([rag_threshold(l) for l in range(min(3, int(self.layout.L)))] + [2.0] * (3 - min(3, int(self.layout.L)))))
        self._xstar_host = (C.c_float * 3)(*([float(xstar_for(rag_threshold(l))) for l in range(min(3, int(self.layout.L)))] + [0.0] * (3 - min(3, int(self.layout.L)))))
        self._filter_scratch = None
        # per-level copies of the index rows (pitch rounded to 4 floats): 48 MB for 1M x 1536, L2 resident
        self._lvl_rows, pitches = [], []
        for l in range(int(self.layout.L)):
            off, keff = int(self.layout.lvl_off[l]), int(self.layout.lvl_keff[l])
            pitch = (keff + 3) // 4 * 4
            self._lvl_rows.append(self.idx[:, off:off + pitch].contiguous())
            pitches.append(pitch)
        self._lvl_ptrs = (C.c_void_p * 3)(*([t.data_ptr() for t in self._lvl_rows[:3]] + [None] * (3 - min(3, len(self._lvl_rows)))))
        self._lvl_pitch = (C.c_int32 * 3)(*(pitches[:3] + [0] * (3 - min(3, len(pitches)))))
        # tensor-core threshold pass: tf32 hi/lo-split, norm-scaled copy of the index rows [N, 128] + validity words
        self.tc_packed = self.tc_valid = self.tc_rows = None
        self.tc_valid_pitch = 0
        if tc_filter and self.fast_filter_ok and self.N > 0 and bool(lib.hq_filter_tc_supported(C.byref(self.layout))):
            self.tc_valid_pitch = int(lib.hq_filter_tc_valid_pitch(self.N))
            cols = int(lib.hq_filter_tc_packed_cols(C.byref(self.layout)))       # [hi | lo] per level, whole 128-byte slabs
            self.tc_packed = torch.empty((self.N, cols), dtype=torch.float32, device=d)
            self.tc_valid = torch.empty((int(self.layout.L), self.tc_valid_pitch), dtype=torch.int32, device=d)
            with torch.cuda.device(d):
                check(lib.hq_filter_tc_pack(dev.ptr(self.idx), dev.ptr(self.level_norms), self.N, C.byref(self.layout), 0,
                                            dev.ptr(self.tc_packed), dev.stream_ptr()))
                check(lib.hq_filter_tc_valid(dev.ptr(self.level_norms), self.N, C.byref(self.layout), dev.ptr(self.tc_valid),
                                             self.tc_valid_pitch, dev.stream_ptr()))
            if self.exc_rows.numel():                       # exceptional rows never pass in the tensor-core pass
                r = self.exc_rows.to(torch.int64)
                keep = ~torch.bitwise_left_shift(torch.ones_like(r), r & 31).to(torch.int32)
                for w, m in zip((r >> 5).tolist(), keep.tolist()):     # a handful of rows: per-word read-modify-write
                    self.tc_valid[:, w] &= m
            # latency path (a handful of queries): the used part of the index rows as plain fp32, scaled by 1 / |c_l|
            # (144 bytes per 1536-D row) -- the window pass of such a batch runs over them on the CUDA cores
            rcols = int(lib.hq_filter_rows_cols(C.byref(self.layout)))
            if rcols and self.N * rcols * 4 <= ROW_PASS_MAX_BYTES:
                self.tc_rows = torch.empty((self.N, rcols), dtype=torch.float32, device=d)
                with torch.cuda.device(d):
                    check(lib.hq_filter_rows_pack(dev.ptr(self.idx), dev.ptr(self.level_norms), self.N, C.byref(self.layout),
                                                  dev.ptr(self.tc_rows), dev.stream_ptr()))

    @property
    def num_levels(self) -> int:
        return int(self.layout.L)


def to_bf16(x: torch.Tensor, norms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """float32 [N, D] -> bf16 [N, pitch] (round to nearest even, pitch = D rounded up to 8, zero padded); with `norms`
    the rows are divided by their norm first (unit rows, zero rows stay zero)."""
    N, D = x.shape
    pitch = (D + 7) // 8 * 8
    out = torch.empty((N, pitch), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        if norms is None:
            check(lib.hq_to_bf16(dev.ptr(x), N, D, x.stride(0) if N else D, dev.ptr(out), pitch, dev.stream_ptr()))
        else:
            check(lib.hq_to_bf16_unit(dev.ptr(x), N, D, x.stride(0) if N else D, dev.ptr(norms), dev.ptr(out), pitch, dev.stream_ptr()))
    return out


_INGEST_PLANS: dict = {}


def ingest_plan_codes(n: int) -> Optional[np.ndarray]:
    """The variant-C compact plan of an n x n grid as (pyramid level << 24) | position-in-curve-order per index slot, or
    None when a slot reads raw cells or a level below 3 / above 6 (hq_shard_ingest reduces levels 3..6 only)."""
    plan, _, _ = plans.c_plan(n, "compact")
    cells = n * n
    codes = np.empty(len(plan), dtype=np.int32)
    for s, e in enumerate(plan.tolist()):
        if e < cells:
            return None
        off, k = e - cells, 1
        while off >= cells >> (2 * k):
            off -= cells >> (2 * k)
            k += 1
        if k < 3 or k > 6:
            return None
        codes[s] = (k << 24) | off
    return codes


def shard_ingest(emb: torch.Tensor, n: int, want_bf16: bool = True):
    """Index rows (variant C, compact), row norms and bf16 unit rows of a block of embeddings in ONE pass
    (hq_shard_ingest); None when the shape is not covered (the caller then runs the three separate passes, whose
    results are bit identical)."""
    N, D = emb.shape
    d = emb.device
    if N == 0 or emb.stride(1) != 1 or emb.stride(0) % 4 or emb.data_ptr() % 16 or D > n * n \
            or not lib.hq_shard_ingest_supported(D):
        return None
    key = (n, d)
    if key not in _INGEST_PLANS:
        codes = ingest_plan_codes(n)
        _INGEST_PLANS[key] = None if codes is None else torch.from_numpy(codes).to(d)
    codes = _INGEST_PLANS[key]
    if codes is None:
        return None
    Lsum = int(codes.numel())
    idx = torch.empty((N, Lsum), dtype=torch.float32, device=d)
    norms = torch.empty(N, dtype=torch.float32, device=d)
    unit = torch.empty((N, D), dtype=torch.bfloat16, device=d) if want_bf16 else None
    with torch.cuda.device(d):
        emb_stride = emb.stride(0) if N > 1 else D           # (a size-1 axis may report any stride)
        check(lib.hq_shard_ingest(dev.ptr(emb), N, D, emb_stride, dev.ptr(codes), Lsum, dev.ptr(idx), idx.stride(0),
                                  dev.ptr(norms), dev.ptr(unit) if unit is not None else None, D, dev.stream_ptr()))
    return idx, norms, unit


def row_lengths(idx: torch.Tensor, layout: IndexLayout) -> torch.Tensor:
    N = idx.shape[0]
    lens = torch.empty((N, int(layout.L)), dtype=torch.int16, device=idx.device)      # uint16 payload
    with torch.cuda.device(idx.device):
        check(lib.hq_index_row_lengths(dev.ptr(idx), N, C.byref(layout), dev.ptr(lens), dev.stream_ptr()))
    return lens


def row_norms(x: torch.Tensor) -> torch.Tensor:
    N, D = x.shape
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.hq_row_norms(dev.ptr(x), N, D, x.stride(0) if N else D, dev.ptr(out), dev.stream_ptr()))
    return out


@dataclass
class FilterTrace:
    """Per-level survivor counts of one filter call (for tests and diagnostics)."""
    n_alive: List[torch.Tensor]
    n_pass: List[torch.Tensor]
    n_out: List[torch.Tensor]


def _mask_words(N: int) -> int:
    return (N + 31) // 32


def progressive_filter(db: EmbeddingDatabase, q_idx: torch.Tensor, q_lens: torch.Tensor, scores: torch.Tensor,
                       mask: torch.Tensor, trace: Optional[FilterTrace] = None, keep_level_scores: Optional[list] = None,
                       tie_rule: str = "id"):
    """Run all filter levels for the queries of one chunk.  `scores` [Qc, N] float32 and
    `mask` [Qc, words] int32 are caller-provided work buffers; on return `mask` holds the
    survivor bits and `scores` the last level's scores (-1 for rows dead before it).
    tie_rule="reference": rows tied exactly at a ratio cut keep the previous levels' order (the reference's stable sort,
    rag/search/engine.py:236) instead of the row id order; costs a copy of the score matrix per level."""
    if tie_rule not in ("id", "reference"):
        raise ValueError("tie_rule must be 'id' or 'reference'")
    prev: list = []
    Qc = q_idx.shape[0]
    N = db.N
    d = db.device
    counts = torch.zeros((db.num_levels, 3, Qc), dtype=torch.int32, device=d)
    with torch.cuda.device(d):
        st = dev.stream_ptr()
        for level in range(db.num_levels):
            n_alive, n_pass, n_out = counts[level, 0], counts[level, 1], counts[level, 2]
            check(lib.hq_filter_level(dev.ptr(db.idx), dev.ptr(db.lens), N, C.byref(db.layout), level,
                                      dev.ptr(q_idx), dev.ptr(q_lens), Qc,
                                      dev.ptr(mask) if level > 0 else None, mask.stride(0),
                                      rag_threshold(level), dev.ptr(scores), scores.stride(0), dev.ptr(mask),
                                      dev.ptr(n_alive), dev.ptr(n_pass), st))
            if tie_rule == "reference" and prev:
                p1 = prev[-1]
                p0 = prev[-2] if len(prev) > 1 else None
                check(lib.hq_filter_select_prev(dev.ptr(scores), scores.stride(0), N, Qc, dev.ptr(p1), dev.ptr(p0), p1.stride(0),
                                                dev.ptr(n_alive), dev.ptr(n_pass), rag_ratio(level), dev.ptr(mask), mask.stride(0),
                                                dev.ptr(n_out), st))
            else:
                check(lib.hq_filter_select(dev.ptr(scores), scores.stride(0), N, Qc, dev.ptr(n_alive), dev.ptr(n_pass),
                                           rag_ratio(level), dev.ptr(mask), mask.stride(0), dev.ptr(n_out), st))
            if tie_rule == "reference" and level + 1 < db.num_levels:
                prev.append(scores[:Qc].clone())
                prev = prev[-2:]
            if keep_level_scores is not None:
                keep_level_scores.append(scores.clone())
    if trace is not None:
        for level in range(db.num_levels):
            trace.n_alive.append(counts[level, 0])
            trace.n_pass.append(counts[level, 1])
            trace.n_out.append(counts[level, 2])
    return mask


def _unpack_bits(mask: torch.Tensor, N: int) -> torch.Tensor:
    """[Q, words] int32 bit mask -> bool [Q, N] on the same device."""
    sh = torch.arange(32, device=mask.device, dtype=torch.int32)
    return ((mask.unsqueeze(-1) >> sh) & 1).to(torch.bool).reshape(mask.shape[0], -1)[:, :N]


def _pack_bits(keep: torch.Tensor, words: int) -> torch.Tensor:
    """bool [Q, N] -> [Q, words] int32 bit mask."""
    Q, N = keep.shape
    pad = torch.zeros((Q, words * 32), dtype=torch.int64, device=keep.device)
    pad[:, :N] = keep
    sh = torch.arange(32, device=keep.device, dtype=torch.int64)
    w = (pad.view(Q, words, 32) << sh).sum(-1)
    return torch.where(w >= (1 << 31), w - (1 << 32), w).to(torch.int32)


def progressive_filter_global(db: EmbeddingDatabase, q_idx: torch.Tensor, q_lens: torch.Tensor, scores: torch.Tensor,
                              mask: torch.Tensor, group=None, trace: Optional[FilterTrace] = None):
    """`filter_scope="global"`: level scores and threshold tests by hq_filter_level on this shard, the ratio
    cut over the GLOBAL candidate list of each query (distributed.global_ratio_cut)."""
    from .distributed import global_ratio_cut, global_ratio_cut_device
    Qc, N, d = q_idx.shape[0], db.N, db.device
    words = _mask_words(N)
    with torch.cuda.device(d):
        st = dev.stream_ptr()
        for level in range(db.num_levels):
            n_alive = torch.zeros(Qc, dtype=torch.int32, device=d)
            n_pass = torch.zeros(Qc, dtype=torch.int32, device=d)
            # the ratio cut over the global list on the device (distributed.global_ratio_cut_device: histograms of the score
            # bits, no [Q, N] temporaries); an empty shard contributes nothing but takes part in every collective
            if N > 0:
                check(lib.hq_filter_level(dev.ptr(db.idx), dev.ptr(db.lens), N, C.byref(db.layout), level,
                                          dev.ptr(q_idx), dev.ptr(q_lens), Qc,
                                          dev.ptr(mask) if level > 0 else None, mask.stride(0),
                                          rag_threshold(level), dev.ptr(scores), scores.stride(0), dev.ptr(mask),
                                          dev.ptr(n_alive), dev.ptr(n_pass), st))
            if os.environ.get("HQ_GLOBAL_CUT_TORCH"):     # the first implementation (eager PyTorch on [Q, N] tensors), kept for A/B
                if N == 0:
                    none = torch.zeros((Qc, 0), dtype=torch.bool, device=d)
                    _, n_out = global_ratio_cut(torch.zeros((Qc, 0), dtype=torch.float32, device=d), none, n_alive, rag_ratio(level),
                                                db.id_base, group)
                else:
                    passed = _unpack_bits(mask[:Qc, :words], N)
                    keep, n_out = global_ratio_cut(scores[:Qc, :N], passed, n_alive, rag_ratio(level), db.id_base, group)
                    mask[:Qc, :words] = _pack_bits(keep, words)
            else:
                n_out = global_ratio_cut_device(scores, mask, N, n_alive, n_pass, rag_ratio(level), db.id_base, group)
            if trace is not None:
                trace.n_alive.append(n_alive)
                trace.n_pass.append(n_pass)
                trace.n_out.append(n_out)
    return mask


def progressive_filter_fast(db: EmbeddingDatabase, q_idx: torch.Tensor, mask: torch.Tensor,
                            trace: Optional[FilterTrace] = None, tensor_cores: bool = True, row_pass: bool = True):
    """All filter levels for a query batch through hq_filter_fast (no score matrix).  With
    `tensor_cores` (and a packed operand on the shard) the threshold pass runs on tcgen05; a handful of queries
    (`row_pass`, hq_filter_rows_max_queries) run their window pass over the shard's scaled fp32 rows instead."""
    use_tc = tensor_cores and db.tc_packed is not None
    Q, N, d = q_idx.shape[0], db.N, db.device
    L = db.num_levels
    need = int(lib.hq_filter_fast_scratch_bytes(N, Q, C.byref(db.layout)))
    if db._filter_scratch is None or db._filter_scratch.numel() < need:
        db._filter_scratch = torch.empty(need, dtype=torch.uint8, device=d)
    n_out = torch.empty(Q, dtype=torch.int32, device=d)
    counts = torch.zeros((L, 3, Q), dtype=torch.int32, device=d) if trace is not None else None
    with torch.cuda.device(d):
        check(lib.hq_filter_fast_rows(dev.ptr(db.idx), dev.ptr(db.level_norms), N, C.byref(db.layout), dev.ptr(q_idx), Q,
                                 C.cast(db._xstar_host, C.c_void_p), C.cast(db._ratio_host, C.c_void_p),
                                 C.cast(db._lvl_ptrs, C.c_void_p), C.cast(db._lvl_pitch, C.c_void_p),
                                 dev.ptr(db.tc_packed) if use_tc else None,
                                 dev.ptr(db.tc_rows) if (use_tc and db.tc_rows is not None and row_pass) else None,
                                 dev.ptr(db.tc_valid) if use_tc else None,
                                 db.tc_valid_pitch if use_tc else 0,
                                 dev.ptr(db.lens), dev.ptr(db.exc_rows) if db.exc_rows.numel() else None, int(db.exc_rows.numel()),
                                 C.cast(db._thr_host, C.c_void_p),
                                 dev.ptr(mask), mask.stride(0), dev.ptr(n_out), dev.ptr(counts),
                                 dev.ptr(db._filter_scratch), db._filter_scratch.numel(), dev.stream_ptr()))
    if trace is not None:
        for level in range(L):
            trace.n_alive.append(counts[level, 0])
            trace.n_pass.append(counts[level, 1])
            trace.n_out.append(counts[level, 2])
    return mask


def prepare_queries(db: EmbeddingDatabase, queries) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """queries [Q, D] -> (q float32 on device, q_idx compact rows, q_lens, q_norms)."""
    q = dev.f32_device(queries, db.device)
    if q.dim() == 1:
        q = q.reshape(1, -1)
    if q.shape[1] != db.D:
        raise ValueError(f"query dimension {q.shape[1]} does not match database dimension {db.D}")
    fused = shard_ingest(q, db.n, want_bf16=False)          # index rows + norms in one launch (bit identical to the two below)
    if fused is not None:
        q_idx, q_norms, _ = fused
        return q, q_idx, row_lengths(q_idx, db.layout), q_norms
    _, q_idx = map_and_index(q, db.n, variant="C", layout="compact", want_grid=False)
    return q, q_idx, row_lengths(q_idx, db.layout), row_norms(q)


ROW_PASS_MAX_BYTES = 4 << 30        # largest scaled-row copy a shard keeps for the CUDA-core window pass of small batches
SPARSE_RERANK_MAX_QUERIES = 4       # batches up to this size score only the filter's survivors (exact fp32, one warp per row)
_SIDE_STREAMS: dict = {}
_FLAG_POOL: list = []          # pinned one-byte buffers (allocating pinned memory per call would synchronise the device)


class _DenseCheck:
    """`(q_lens == keff).all()` evaluated on the main stream, read back through a side stream so that the host can
    wait for it without waiting for kernels launched afterwards on the main stream."""

    def __init__(self, q_lens: torch.Tensor, db: "EmbeddingDatabase"):
        d = q_lens.device
        flag = (q_lens == db._keff).all()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(d))
        side = _SIDE_STREAMS.get(d)
        if side is None:
            side = _SIDE_STREAMS[d] = torch.cuda.Stream(device=d, priority=-1)
        try:
            self.host = _FLAG_POOL.pop()
        except IndexError:
            self.host = torch.empty(1, dtype=torch.bool).pin_memory()
        self.done = torch.cuda.Event()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            self.host.copy_(flag.reshape(1), non_blocking=True)
            self.done.record(side)
        flag.record_stream(side)

    def result(self) -> bool:
        self.done.synchronize()
        ok = bool(self.host[0])
        _FLAG_POOL.append(self.host)
        return ok


def packed_result_buffers(Q: int, k: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """ids int64 [Q, k] and scores float32 [Q, k] carved out of ONE allocation (ids first), so that a row-sharded
    search can ship both with a single all-gather (`distributed.allgather_merge`)."""
    buf = torch.empty(3 * Q * k, dtype=torch.int32, device=device)
    ids = buf[: 2 * Q * k].view(torch.int64).view(Q, k)
    scores = buf[2 * Q * k:].view(torch.float32).view(Q, k)
    return ids, scores


def search_batch(db: EmbeddingDatabase, queries, k: int = 10, *, use_filter: bool = True,
                 work_bytes: int = 4 << 30, return_mask: bool = False, trace: Optional[FilterTrace] = None,
                 rerank: str = "auto", filter_impl: str = "auto", filter_scope: str = "shard", group=None,
                 filter_scratch_bytes: int = 24 << 30, _dense_flag: Optional[list] = None, guard_stats: Optional[list] = None,
                 tie_rule: str = "id"):
    """Progressive top-k of a batch of query embeddings against one shard.

    `tie_rule`: rows whose level score ties EXACTLY at a ratio cut are kept by ascending row id ("id", every filter
    path) or in the order of the previous levels' scores like the reference's stable sort ("reference",
    rag/search/engine.py:236; exact filter path only -- it is selected automatically).

    Returns (ids int64 [Q, k] (-1 = fewer than k survivors), scores float32 [Q, k]).  Scores are
    (cos + 1) / 2 of the full embeddings; ties resolve to the lower row id.
    `filter_scope="global"` (row-sharded databases): the ratio cut of every filter level ranks the
    candidates of ALL shards of `group` together, like the reference's single list.
    `guard_stats`: a list that receives one int32 device tensor per tensor-core rerank launch ([0] = queries whose
    bf16 shortlist could not be proven sufficient and were re-scored exactly over all survivors, [1] = rows re-scored)."""
    tok = _phase("query_index")
    q, q_idx, q_lens, q_norms = prepare_queries(db, queries)
    _end(tok)
    Q, N, d = q.shape[0], db.N, db.device
    ids, out_scores = packed_result_buffers(Q, k, d)
    if filter_scope not in ("shard", "global"):
        raise ValueError("filter_scope must be 'shard' or 'global'")
    collective = filter_scope == "global" and use_filter          # every rank must take part in the per-level collectives,
    if Q == 0 or (N == 0 and not collective):                     # an empty shard included
        ids.fill_(-1)
        out_scores.fill_(-1.0)
        return (ids, out_scores, None) if return_mask else (ids, out_scores)
    if rerank == "auto":
        rerank = "bf16" if (db.emb_bf16 is not None and k <= 20) else "f32"
        # A handful of queries: the filter's ratio cuts keep at most 0.3 * 0.5 * 0.7 = 10.5 % of the rows, so scoring only
        # the survivors in fp32 (<= Q * 0.105 * N rows of 4 D bytes) reads less than the dense pass over the bf16 copy
        # (N rows of 2 D bytes) up to Q = 4.  Same scores: the tensor-core path re-scores its shortlist with this arithmetic.
        if use_filter and Q <= SPARSE_RERANK_MAX_QUERIES and db.emb is not None:
            rerank = "sparse"
    if rerank not in ("bf16", "f32", "sparse", "sparse_bf16"):
        raise ValueError("rerank must be 'auto', 'bf16', 'f32', 'sparse' or 'sparse_bf16'")
    # "sparse_bf16": the survivors are ranked by their bf16 unit rows first (exact re-score of the best 32, proven or redone
    # over the fp32 rows: same results).  Not the default: scattered rows cost per row, not per byte -- 86 us against 90 us
    # for a query against 1 M x 1536, and one more launch.
    sparse_bf16 = rerank == "sparse_bf16" and db.emb_bf16 is not None
    if rerank == "sparse_bf16":
        rerank = "sparse"
    if rerank == "bf16" and (db.emb_bf16 is None or k > 20):
        raise ValueError("the tensor-core rerank needs a bf16 database copy and k <= 20")
    if rerank != "bf16" and db.emb is None:
        raise ValueError("a bf16-only database (EmbeddingDatabase.from_chunks) is searched by the tensor-core rerank only (k <= 20)")
    q_bf16 = None
    if rerank == "bf16":
        tok = _phase("query_index")
        q_bf16 = to_bf16(q)
        _end(tok)
    words = _mask_words(N)
    if filter_impl not in ("auto", "fast", "fast_fp32", "exact"):
        raise ValueError("filter_impl must be 'auto', 'fast', 'fast_fp32' or 'exact'")
    n_chunk_rows = N
    if tie_rule not in ("id", "reference"):
        raise ValueError("tie_rule must be 'id' or 'reference'")
    if tie_rule == "reference":
        if filter_scope == "global" or filter_impl in ("fast", "fast_fp32"):
            raise ValueError("tie_rule='reference' is implemented by the exact shard-local filter only")
        filter_impl = "exact"
    if filter_scope == "global":
        filter_impl = "exact"                      # the global cut works on the per-level score matrix
        if collective:
            # the number of query chunks (= collectives per level) must be the same on every rank: size the chunks by the
            # LARGEST shard (shard_bounds gives base or base + 1 rows, and a shard may be empty)
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(group) > 1:
                nmax = torch.tensor([N], dtype=torch.int64, device=d)
                dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
                n_chunk_rows = int(nmax.item())
    fast = False
    dense_queries = None
    if use_filter and filter_impl != "exact":
        fast = db.fast_filter_ok
        if filter_impl in ("fast", "fast_fp32") and not fast:
            raise ValueError("the fast filter needs dense index rows (all stored lengths structural) and L <= 3")
        if fast:
            # The fast path also needs dense QUERY index rows.  The test runs on the device and its result travels to
            # the host on a SIDE stream that only waits for the query-side kernels; the filter and the rerank are
            # launched speculatively on the main stream and the host looks at the flag after the launches.  The host
            # therefore never waits for the big kernels of this batch (a read-back on the main stream did: ~0.3 ms of
            # idle GPU per batch while the host prepared the next one); a batch with a sparse query is searched
            # again through the exact path.
            if _dense_flag is not None:             # graph capture (SearchGraph): the caller reads the flag after the replay
                _dense_flag.append((q_lens == db._keff).all())
            else:
                dense_queries = _DenseCheck(q_lens, db)
    if fast or not use_filter:
        work_bytes = max(work_bytes, 4 * N * Q) if rerank == "bf16" else work_bytes
    qc = int(max(1, min(Q, work_bytes // (4 * max(1, n_chunk_rows))))) if not (rerank == "bf16" and (fast or not use_filter)) else Q
    if fast and use_filter:
        # the fast filter keeps bit planes and candidate lists for the whole query chunk in scratch: bound it
        # (C5: 12.5 M rows x 4096 queries would ask for > 100 GB) by searching the batch in 128-query multiples
        budget = max(int(filter_scratch_bytes), 1 << 28)
        while qc > 128 and int(lib.hq_filter_fast_scratch_bytes(N, qc, C.byref(db.layout))) > budget:
            qc = max(128, (qc // 2 + 127) // 128 * 128)
    # a handful of queries behind the filter: rerank + top-k of the survivors in one launch, no [Q, N] score row
    fused_sparse = (rerank == "sparse" and use_filter and N > 0 and
                    bool(lib.hq_rerank_sparse_topk_supported(db.D, db.emb.stride(0), q.stride(0), k)))
    need_scores = not ((rerank == "bf16" and (fast or not use_filter)) or (fused_sparse and fast))    # the exact filter's scratch
    scores = torch.empty((qc, N), dtype=torch.float32, device=d) if need_scores else None
    # rows padded to whole 32-byte sectors: the window pass of the fast filter then writes its plane straight into the mask
    mask = torch.zeros((qc, (words + 7) // 8 * 8), dtype=torch.int32, device=d)[:, :words]
    masks = [] if return_mask else None
    with torch.cuda.device(d):
        for s in range(0, Q, qc):
            e = min(Q, s + qc)
            nq = e - s
            m = None
            if N == 0:                              # empty shard in a global-scope search: collectives only, no results
                progressive_filter_global(db, q_idx[s:e], q_lens[s:e], scores[:nq], mask[:nq], group, trace)
                ids[s:e].fill_(-1)
                out_scores[s:e].fill_(-1.0)
                continue
            if use_filter:
                tok = _phase("filter")
                if fast:
                    m = progressive_filter_fast(db, q_idx[s:e], mask[:nq], trace, tensor_cores=filter_impl != "fast_fp32")
                elif filter_scope == "global":
                    m = progressive_filter_global(db, q_idx[s:e], q_lens[s:e], scores[:nq], mask[:nq], group, trace)
                else:
                    m = progressive_filter(db, q_idx[s:e], q_lens[s:e], scores[:nq], mask[:nq], trace, tie_rule=tie_rule)
                _end(tok)
                if return_mask:
                    masks.append(m.clone())
            if rerank == "bf16":
                tok = _phase("rerank_gemm")
                sb = int(lib.hq_rerank_bf16_scratch_bytes(N, nq, k))
                scratch = torch.empty(sb, dtype=torch.uint8, device=d)
                guard = torch.empty(4 + nq, dtype=torch.int32, device=d)   # [0] queries sent to the exact fallback, [1] rows re-scored
                check(lib.hq_rerank_topk_unit_bf16(dev.ptr(db.emb_bf16), db.emb_bf16.stride(0), dev.ptr(db.emb),
                                              db.emb.stride(0) if db.emb is not None else 0,
                                              dev.ptr(db.norms), dev.ptr(db.zero_rows), int(db.zero_rows.numel()), N, db.D,
                                              dev.ptr(q_bf16[s:e]), q_bf16.stride(0),
                                              dev.ptr(q[s:e]), q.stride(0), dev.ptr(q_norms[s:e]), nq,
                                              dev.ptr(m), mask.stride(0), k, db.id_base, float(db.dc_max),
                                              dev.ptr(ids[s:e]), dev.ptr(out_scores[s:e]), dev.ptr(guard), dev.ptr(scratch), sb,
                                              dev.stream_ptr()))
                if guard_stats is not None:
                    guard_stats.append(guard)
                _end(tok)
                continue
            if fused_sparse and m is not None:
                tok = _phase("rerank_gemm")
                sb = int(lib.hq_rerank_sparse_topk_scratch_bytes(nq, k))
                scratch = torch.empty(sb, dtype=torch.uint8, device=d)
                b16 = db.emb_bf16 if sparse_bf16 else None
                check(lib.hq_rerank_sparse_topk(dev.ptr(db.emb), dev.ptr(db.norms), N, db.D, db.emb.stride(0),
                                                    dev.ptr(b16), b16.stride(0) if b16 is not None else 0, float(db.dc_max),
                                                    dev.ptr(q[s:e]), dev.ptr(q_norms[s:e]), nq, q.stride(0),
                                                    dev.ptr(m), mask.stride(0), k, db.id_base,
                                                    dev.ptr(ids[s:e]), dev.ptr(out_scores[s:e]), dev.ptr(scratch), sb, dev.stream_ptr()))
                _end(tok)
                continue
            tok = _phase("rerank_gemm")
            score_fn = lib.hq_rerank_scores_sparse_f32 if (rerank == "sparse" and m is not None) else lib.hq_rerank_scores_f32
            check(score_fn(dev.ptr(db.emb), dev.ptr(db.norms), N, db.D, db.emb.stride(0),
                                           dev.ptr(q[s:e]), dev.ptr(q_norms[s:e]), nq, q.stride(0),
                                           dev.ptr(m), mask.stride(0), dev.ptr(scores), scores.stride(0), dev.stream_ptr()))
            _end(tok)
            tok = _phase("topk")
            tb = int(lib.hq_topk_chunked_scratch_bytes(N, nq, k))
            tscratch = torch.empty(max(tb, 8), dtype=torch.uint8, device=d)
            check(lib.hq_topk_from_scores_chunked(dev.ptr(scores), scores.stride(0), N, nq, k, db.id_base,
                                                  dev.ptr(ids[s:e]), dev.ptr(out_scores[s:e]), dev.ptr(tscratch), tb, dev.stream_ptr()))
            _end(tok)
    if dense_queries is not None and not dense_queries.result():
        # a batch with a sparse query (an exactly-zero block mean at the end of an index row): exact path
        if filter_impl in ("fast", "fast_fp32"):
            raise ValueError("the fast filter needs dense index rows (all stored lengths structural) and L <= 3")
        if trace is not None:
            trace.n_alive.clear(); trace.n_pass.clear(); trace.n_out.clear()
        del mask, scores, ids, out_scores
        return search_batch(db, queries, k, use_filter=use_filter, work_bytes=work_bytes, return_mask=return_mask, trace=trace,
                            rerank=rerank, filter_impl="exact", filter_scope=filter_scope, group=group,
                            filter_scratch_bytes=filter_scratch_bytes, guard_stats=guard_stats, tie_rule=tie_rule)
    if return_mask:
        return ids, out_scores, (torch.cat(masks) if masks else None)
    return ids, out_scores


def search_stream(db: EmbeddingDatabase, host_batches, k: int = 10, *, depth: int = 2, post=None, post_stream=None, **kw):
    """Throughput path for queries that live in HOST memory: a generator over `host_batches` (CPU tensors [Q, D], pinned for
    asynchronous copies; every batch the same shape) that yields `(ids, scores)` as pinned host tensors, in order.  The copy
    of batch i + 1 runs on a copy stream while batch i is searched and the results of batch i - 1 travel back, so the
    host never sits between a synchronisation and the next launch (a copy -> search -> read-back -> synchronise loop leaves the
    GPU idle for the copy and for the host's launch preparation: 20 % of a 1 ms batch on a 125 K-row shard).  `post(ids,
    scores)` runs on the device results before they are read back (the all-gather merge of a row-sharded search).  The
    yielded tensors belong to one of `depth` slots and are overwritten `depth` batches later."""
    from collections import deque
    d = db.device
    main = torch.cuda.current_stream(d)
    # the copy stream and the slots (device query buffer, pinned result buffers, events) live on the shard and are reused by
    # later calls: creating them is not free (a pinned allocation while the GPU is busy stalled the second batch of EVERY
    # call by ~125 ms, 25 batches' worth of search at 1 M rows)
    state = db.__dict__.setdefault("_stream_state", {})
    copy_stream = state.get("copy_stream")
    if copy_stream is None:
        copy_stream = state["copy_stream"] = torch.cuda.Stream(device=d)
    slots: list = []
    pending: deque = deque()

    def finish(slot):
        slot["done"].synchronize()
        return slot["ids"], slot["scores"]

    for i, qh in enumerate(host_batches):
        qh = qh if isinstance(qh, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(qh, dtype=np.float32))
        if qh.dim() == 1:
            qh = qh.reshape(1, -1)
        if i == 0:
            slots = state.setdefault(("slots", tuple(qh.shape), depth), [])
        if len(slots) < depth:
            slots.append({"q": torch.empty(tuple(qh.shape), dtype=torch.float32, device=d), "ids": None, "scores": None,
                          "copied": torch.cuda.Event(), "done": torch.cuda.Event()})
        slot = slots[i % depth]
        if len(pending) == depth:                    # the slot's previous batch: finished before its buffers are reused
            yield finish(pending.popleft())
        if tuple(qh.shape) != tuple(slots[0]["q"].shape):
            raise ValueError(f"search_stream needs batches of one shape, got {tuple(qh.shape)} after {tuple(slots[0]['q'].shape)}")
        with torch.cuda.stream(copy_stream):
            slot["q"].copy_(qh, non_blocking=True)
            slot["copied"].record(copy_stream)
        main.wait_event(slot["copied"])
        ids, sc = search_batch(db, slot["q"], k, **kw)
        if slot["ids"] is None or tuple(slot["ids"].shape) != tuple(ids.shape):
            slot["ids"] = torch.empty(tuple(ids.shape), dtype=ids.dtype).pin_memory()
            slot["scores"] = torch.empty(tuple(sc.shape), dtype=sc.dtype).pin_memory()
        if post is not None and post_stream is not None:
            # the post step (the all-gather merge of a row-sharded search) and the read-back run on `post_stream`, ordered
            # after this batch's search; the main stream goes straight on to the next batch
            searched = torch.cuda.Event()
            searched.record(main)
            post_stream.wait_event(searched)
            with torch.cuda.stream(post_stream):
                out_i, out_s = post(ids, sc)
                for t in (ids, sc, out_i, out_s):
                    t.record_stream(post_stream)
                slot["ids"].copy_(out_i, non_blocking=True)
                slot["scores"].copy_(out_s, non_blocking=True)
                slot["done"].record(post_stream)
        else:
            if post is not None:
                ids, sc = post(ids, sc)
            slot["ids"].copy_(ids, non_blocking=True)
            slot["scores"].copy_(sc, non_blocking=True)
            slot["done"].record(main)
        pending.append(slot)
    while pending:
        yield finish(pending.popleft())


class SearchGraph:
    """`search_batch` for a fixed (shard, batch size, k) captured ONCE as a CUDA graph and replayed per batch.

    A small batch is launch bound: a single query against 1 M x 1536 runs ~0.45 ms of kernels (threshold pass, list
    cascade, sparse rerank, two-level top-k and a dozen query-side kernels) but took 0.67 ms per call, the rest being the
    host preparing fifteen launches.  The replay issues them as one graph launch.  Results are identical to
    `search_batch` (same kernels, same order).  The dense-query test that `search_batch` reads back through a side stream
    is part of the graph; `search` looks at the flag together with the results and sends a batch with a sparse query
    index row through the exact path, like `search_batch` does."""

    def __init__(self, db: EmbeddingDatabase, batch: int, k: int = 10, **kw):
        if any(key in kw for key in ("return_mask", "trace", "group")) or kw.get("filter_scope", "shard") != "shard":
            raise ValueError("SearchGraph captures the plain shard-local search (no mask / trace / global scope)")
        self.db, self.Q, self.k, self.kw = db, int(batch), int(k), dict(kw)
        d = db.device
        self.q = torch.empty((self.Q, db.D), dtype=torch.float32, device=d)
        if db.N > 0:
            rows = torch.arange(self.Q, device=d) % db.N                  # warm-up input with dense index rows
            self.q.copy_(db.emb[rows] if db.emb is not None else db.emb_bf16[rows, : db.D].to(torch.float32))
        else:
            self.q.fill_(1.0)
        side = torch.cuda.Stream(device=d)
        side.wait_stream(torch.cuda.current_stream(d))
        with torch.cuda.stream(side):                # eager warm-up: plans, scratch and kernel attributes outside the graph
            for _ in range(2):
                search_batch(db, self.q, self.k, _dense_flag=[], **self.kw)
        torch.cuda.current_stream(d).wait_stream(side)
        torch.cuda.synchronize(d)
        self.graph = torch.cuda.CUDAGraph()
        flag: list = []
        with torch.cuda.graph(self.graph):
            self.ids, self.scores = search_batch(db, self.q, self.k, _dense_flag=flag, **self.kw)
            self.dense = flag[0].to(torch.int32) if flag else None
        self._dense_host = torch.empty(1, dtype=torch.int32).pin_memory() if self.dense is not None else None
        # the captured kernels point into the shard's filter scratch: keep THAT allocation alive even if a later, larger
        # eager batch makes the shard replace it
        self._keep = (db._filter_scratch,)

    def search(self, queries, sync: bool = True):
        """queries [batch, D] (host or device).  Returns (ids, scores): the graph's output buffers, overwritten by the
        next call.  With sync=False the dense-query flag is not looked at (the caller knows its queries are dense)."""
        q = queries if isinstance(queries, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        if q.dim() == 1:
            q = q.reshape(1, -1)
        if tuple(q.shape) != tuple(self.q.shape):
            raise ValueError(f"this graph searches batches of shape {tuple(self.q.shape)}, got {tuple(q.shape)}")
        self.q.copy_(q, non_blocking=True)
        self.graph.replay()
        if sync and self.dense is not None:
            self._dense_host.copy_(self.dense, non_blocking=True)
            torch.cuda.current_stream(self.db.device).synchronize()
            if int(self._dense_host[0]) == 0:
                return search_batch(self.db, queries, self.k, **{**self.kw, "filter_impl": "exact"})
        return self.ids, self.scores


def unpack_mask(mask: torch.Tensor, N: int) -> np.ndarray:
    """[Q, words] int32 bit mask -> bool [Q, N] on the host."""
    m = mask.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :N].astype(bool)


def cosine01(a, b, device=None) -> float:
    """(cos(a, b) + 1) / 2 of two flattened arrays on their common prefix; 0 if a norm is 0
    (rag/search/engine.py:622-660, :1025-1051)."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    b = np.asarray(b, dtype=np.float32).reshape(-1)
    m = min(a.size, b.size)
    if m == 0:
        return 0.0
    d = dev.require_cuda(device)
    ta = dev.f32_device(a[:m].reshape(1, m), d)
    tb = dev.f32_device(b[:m].reshape(1, m), d)
    out = torch.empty((1, 1), dtype=torch.float32, device=d)
    nb, na = row_norms(tb), row_norms(ta)          # keep both alive until the kernel has run
    with torch.cuda.device(d):
        check(lib.hq_rerank_scores_f32(dev.ptr(tb), dev.ptr(nb), 1, m, m, dev.ptr(ta), dev.ptr(na),
                                       1, m, None, 0, dev.ptr(out), 1, dev.stream_ptr()))
    # |cos| <= 1: a score above 1.0 is rounding noise of dot / (|a| |b|) (identical vectors give 1 + 1 ulp for some inputs;
    # the reference's tests compare such a pair with == 1.0, tests/test_similarity_calculation.py:150-158)
    return min(1.0, float(out.item()))


def granularity_weights(num_levels: int) -> np.ndarray:
    """rag/search/engine.py:1101-1138: 8^(L-i-1), normalised, first level doubled, renormalised."""
    if num_levels <= 0:
        return np.array([])
    if num_levels == 1:
        return np.array([1.0])
    w = np.array([8.0 ** (num_levels - i - 1) for i in range(num_levels)])
    w = w / w.sum()
    w[0] *= 2.0
    return w / w.sum()


def comprehensive_scores(frames, query_frames, original_height: Optional[int] = None, cand_ids=None, device=None) -> torch.Tensor:
    """Comprehensive similarity (0.5 hierarchical + 0.3 cosine + 0.2 spatial locality,
    rag/search/engine.py:516-575) of Q query frames against enhanced frames [N, H + L, W] on the device.
    `cand_ids` [Q, M] (int64, -1 = none) restricts every query to a shortlist.  Returns float32 [Q, M or N]."""
    d = dev.require_cuda(device if device is not None else (frames.device if isinstance(frames, torch.Tensor) and frames.is_cuda else None))
    f = dev.f32_device(frames, d)
    qf = dev.f32_device(query_frames, d)
    if qf.dim() == 2:
        qf = qf.unsqueeze(0)
    if f.dim() != 3 or qf.dim() != 3 or f.shape[1:] != qf.shape[1:]:
        raise ValueError("frames must be [N, H + L, W] and query frames [Q, H + L, W] of the same frame shape")
    N, rows, W = f.shape
    H = int(original_height) if original_height is not None else rows - len(plans.c_levels(W))
    if H != W:
        raise NotImplementedError("comprehensive similarity is implemented for square grids (original_height == width)")
    L = rows - H
    Q = qf.shape[0]
    ids = None
    M = N
    if cand_ids is not None:
        ids = cand_ids.to(device=d, dtype=torch.int64).contiguous() if isinstance(cand_ids, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(cand_ids, dtype=np.int64)).to(d)
        if ids.dim() != 2 or ids.shape[0] != Q:
            raise ValueError("cand_ids must be [Q, M]")
        M = ids.shape[1]
    out = torch.empty((Q, M), dtype=torch.float32, device=d)
    w = (C.c_float * 8)(*([float(x) for x in granularity_weights(L)] + [0.0] * (8 - L)))
    with torch.cuda.device(d):
        check(lib.hq_comprehensive_scores(dev.ptr(f), N, W, L, rows * W, dev.ptr(qf), Q, rows * W, C.cast(w, C.c_void_p),
                                          dev.ptr(ids), M, dev.ptr(out), dev.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------
# RAG engine surface
# ------------------------------------------------------------------------------------------
class RAGSearchEngineImpl:
    """Component methods of rag/search/engine.py:14 on the device.

    `progressive_hierarchical_search(query_frame)` filters the frames returned by
    `_get_all_candidate_embeddings()` (the hook the reference's own tests patch,
    tests/test_progressive_filtering.py:167-169).

    Two ways of locating the index rows of a frame:
      * default (`strict=False`): explicitly, `original_height = frame_height - default_level_count`,
        all levels of all candidates in one launch per level (`hq_filter_level` / `hq_filter_select`);
      * `strict=True` (what `dropin.install()` uses): the reference's own >= 50 %-zeros height
        heuristic (:134-162), trailing-zero stripping and dropping of all-zero rows (:97-132), per
        frame, with the reference's list bookkeeping (stable sort, so ties keep the previous level's
        order, :236) on the host and every level's scores of all current candidates from one device
        launch per distinct prefix length."""

    strict = False

    def __init__(self, config=None, dual_storage=None, device=None, strict: bool = False):
        self.config = config
        self._device = device
        self.strict = bool(strict)

    def _get_all_candidate_embeddings(self) -> List[np.ndarray]:
        return []

    @staticmethod
    def _split(frame: np.ndarray) -> Tuple[int, int]:
        height, width = frame.shape
        L = len(plans.c_levels(width))
        return max(0, height - L), L

    def _detect_original_embedding_height(self, enhanced_embedding: np.ndarray) -> int:
        """rag/search/engine.py:134-162: the first row from the bottom with fewer than 50 % zeros ends the
        embedding (host scalar logic over one frame)."""
        height, width = enhanced_embedding.shape
        if width == 0:
            return height
        dense = (np.count_nonzero(enhanced_embedding == 0, axis=1) / width) < 0.5
        hit = np.nonzero(dense)[0]
        return int(hit[-1]) + 1 if len(hit) else height

    def _extract_hierarchical_indices(self, embedding_with_indices: np.ndarray) -> List[np.ndarray]:
        if embedding_with_indices.ndim != 2:
            return []
        if self.strict:                                   # rag/search/engine.py:97-132
            height = embedding_with_indices.shape[0]
            H = self._detect_original_embedding_height(embedding_with_indices)
            rows = []
            for r in range(H, height):
                row = embedding_with_indices[r, :]
                nz = np.nonzero(row)[0]
                if len(nz) > 0:                           # all-zero rows are dropped: later rows move up a level
                    rows.append(row[: nz[-1] + 1])
            return rows
        H, _ = self._split(embedding_with_indices)
        rows = []
        for r in range(H, embedding_with_indices.shape[0]):
            row = embedding_with_indices[r, :]
            nz = np.nonzero(row)[0]
            rows.append(row[: nz[-1] + 1] if len(nz) > 0 else row[:1])
        return rows

    def _extract_original_embedding(self, enhanced_embedding: np.ndarray) -> np.ndarray:
        if enhanced_embedding.ndim == 1:
            return enhanced_embedding
        if self.strict:                                   # rag/search/engine.py:604-620
            return enhanced_embedding[: self._detect_original_embedding_height(enhanced_embedding), :]
        return enhanced_embedding[: self._split(enhanced_embedding)[0], :]

    # ---- strict mode: the reference's list flow, scores from the device --------------------------------
    def _prefix_scores(self, query_row: np.ndarray, cand_rows: Sequence[Optional[np.ndarray]]) -> np.ndarray:
        """(cos + 1) / 2 of `query_row` against every candidate row on the common prefix (0.0 for a missing or empty
        row, rag/search/engine.py:208-231): one launch per distinct prefix length."""
        out = np.zeros(len(cand_rows), dtype=np.float64)
        by_len: Dict[int, List[int]] = {}
        for i, r in enumerate(cand_rows):
            m = 0 if r is None else min(len(r), len(query_row))
            if m > 0:
                by_len.setdefault(m, []).append(i)
        if not by_len:
            return out
        d = dev.require_cuda(self._device)
        for m, members in by_len.items():
            rows = dev.f32_device(np.stack([np.asarray(cand_rows[i][:m], dtype=np.float32) for i in members]), d)
            q = dev.f32_device(np.asarray(query_row[:m], dtype=np.float32).reshape(1, m), d)
            sc = torch.empty((1, len(members)), dtype=torch.float32, device=d)
            nr, nq = row_norms(rows), row_norms(q)
            with torch.cuda.device(d):
                check(lib.hq_rerank_scores_f32(dev.ptr(rows), dev.ptr(nr), len(members), m, m, dev.ptr(q), dev.ptr(nq),
                                               1, m, None, 0, dev.ptr(sc), len(members), dev.stream_ptr()))
            out[np.asarray(members)] = sc[0].cpu().numpy()
        return out

    def _apply_progressive_threshold(self, candidate_scores: List[Tuple[int, float]], level: int) -> List[int]:
        """rag/search/engine.py:243-287 (host list logic)."""
        if not candidate_scores:
            return []
        threshold = rag_threshold(level)
        max_candidates = max(1, int(len(candidate_scores) * rag_ratio(level)))
        kept: List[int] = []
        for candidate_idx, score in candidate_scores:
            if score >= threshold and len(kept) < max_candidates:
                kept.append(candidate_idx)
        return kept

    def _filter_candidates_at_level(self, query_level_indices: np.ndarray, candidate_embeddings: List[np.ndarray],
                                    current_candidates: List[int], level: int) -> List[int]:
        """rag/search/engine.py:178-241 with the per-candidate cosine loop replaced by one batched device call."""
        if not current_candidates or len(query_level_indices) == 0:
            return current_candidates
        members, rows = [], []
        for candidate_idx in current_candidates:
            if candidate_idx >= len(candidate_embeddings):
                continue
            cand = self._extract_hierarchical_indices(candidate_embeddings[candidate_idx])
            members.append(candidate_idx)
            rows.append(cand[level] if level < len(cand) else None)
        scores = self._prefix_scores(np.asarray(query_level_indices), rows)
        candidate_scores = [(i, float(s)) for i, s in zip(members, scores)]
        candidate_scores.sort(key=lambda x: x[1], reverse=True)            # stable: ties keep the incoming order (:236)
        return self._apply_progressive_threshold(candidate_scores, level)

    def progressive_hierarchical_search(self, query_embedding: np.ndarray) -> List[int]:
        if query_embedding.size == 0:
            return []
        if self.strict:                                   # rag/search/engine.py:51-95
            query_indices = self._extract_hierarchical_indices(query_embedding)
            if len(query_indices) == 0:
                return []
            cands = self._get_all_candidate_embeddings()
            if not cands:
                return []
            candidates = list(range(len(cands)))
            for level in range(len(query_indices)):
                if not candidates:
                    break
                candidates = self._filter_candidates_at_level(query_indices[level], cands, candidates, level)
            return candidates
        if query_embedding.ndim != 2:
            return []
        cands = self._get_all_candidate_embeddings()
        if not cands:
            return []
        H, L = self._split(query_embedding)
        if L == 0 or H <= 0:
            return []
        d = dev.require_cuda(self._device)
        W = query_embedding.shape[1]
        frames = dev.f32_device(np.stack([np.asarray(c, dtype=np.float32) for c in cands]), d)
        qf = dev.f32_device(np.asarray(query_embedding, dtype=np.float32), d)
        lay = IndexLayout()
        lay.L = L
        lay.Lsum = L * W
        for i in range(L):
            lay.lvl_off[i] = i * W
            lay.lvl_w[i] = W
            lay.lvl_keff[i] = W
        if W > 64:
            raise NotImplementedError("frame-based filter supports index rows up to 64 values")
        idx = frames[:, H:, :].reshape(frames.shape[0], L * W).contiguous()
        q_idx = qf[H:, :].reshape(1, L * W).contiguous()
        N = idx.shape[0]
        lens, q_lens = row_lengths(idx, lay), row_lengths(q_idx, lay)

        class _Shard:
            pass
        shard = _Shard()
        shard.N, shard.device, shard.idx, shard.lens, shard.layout, shard.num_levels = N, d, idx, lens, lay, L
        scores = torch.empty((1, N), dtype=torch.float32, device=d)
        mask = torch.zeros((1, _mask_words(N)), dtype=torch.int32, device=d)
        progressive_filter(shard, q_idx, q_lens, scores, mask)
        alive = unpack_mask(mask, N)[0]
        s = scores[0].cpu().numpy()
        ids = np.nonzero(alive)[0]
        # reference order: sorted by last-level score, descending, stable (:236)
        return [int(i) for i in ids[np.argsort(-s[ids], kind="stable")]]

    def _compare_single_level_indices(self, query_indices: np.ndarray, candidate_indices: np.ndarray) -> float:
        if len(query_indices) == 0 or len(candidate_indices) == 0:
            return 0.0
        return cosine01(query_indices, candidate_indices, self._device)

    def _calculate_embedding_cosine_similarity(self, embedding1: np.ndarray, embedding2: np.ndarray) -> float:
        if embedding1.size == 0 or embedding2.size == 0:
            return 0.0
        return cosine01(embedding1, embedding2, self._device)

    def _get_similarity_weights(self) -> Dict[str, float]:
        """rag/search/engine.py:716-727"""
        return {"hierarchical": 0.5, "embedding": 0.3, "spatial": 0.2}

    def _pad_indices_to_length(self, indices: List[np.ndarray], target_length: int) -> np.ndarray:
        """rag/search/engine.py:577-602 (host copy)."""
        if not indices:
            return np.zeros((target_length, 1))
        padded = np.zeros((target_length, max(len(i) for i in indices)))
        for i, row in enumerate(indices):
            if i < target_length:
                padded[i, : len(row)] = row
        return padded

    def _calculate_spatial_locality_similarity(self, embedding1: np.ndarray, embedding2: np.ndarray) -> float:
        """rag/search/engine.py:662-714: mean (cos + 1) / 2 over w x w windows at stride w / 2 of the two original grids.
        All windows of the pair are scored by ONE device launch (window rows as a [windows, w * w] operand)."""
        if embedding1.shape != embedding2.shape or embedding1.ndim != 2:
            return 0.0
        o1, o2 = self._extract_original_embedding(embedding1), self._extract_original_embedding(embedding2)
        if o1.shape != o2.shape:
            return 0.0
        height, width = o1.shape
        w = min(4, height // 4, width // 4)
        if w < 2:
            return self._calculate_embedding_cosine_similarity(o1, o2)
        ii = np.arange(0, height - w + 1, w // 2)
        jj = np.arange(0, width - w + 1, w // 2)
        if len(ii) == 0 or len(jj) == 0:
            return 0.0

        def windows(o):
            v = np.lib.stride_tricks.sliding_window_view(np.asarray(o, dtype=np.float32), (w, w))
            return np.ascontiguousarray(v[ii][:, jj].reshape(len(ii) * len(jj), w * w))
        d = dev.require_cuda(self._device)
        a, b = dev.f32_device(windows(o1), d), dev.f32_device(windows(o2), d)
        nw = a.shape[0]
        sc = torch.empty(nw, dtype=torch.float32, device=d)
        na, nb = row_norms(a), row_norms(b)
        with torch.cuda.device(d):
            check(lib.hq_paired_cosine01(dev.ptr(a), dev.ptr(na), dev.ptr(b), dev.ptr(nb), nw, w * w, w * w, dev.ptr(sc),
                                         dev.stream_ptr()))
        return float(sc.to(torch.float64).mean().item())

    def _calculate_comprehensive_similarity(self, query_embedding: np.ndarray, query_indices, candidate_frame: np.ndarray,
                                            frame_number: int = 0) -> float:
        """rag/search/engine.py:516-575 for one pair."""
        if self.strict:                  # the reference's composition over this class's (device) leaf methods
            candidate_indices = self._extract_hierarchical_indices(candidate_frame)
            hier = 0.0
            if query_indices and candidate_indices:
                levels = max(len(query_indices), len(candidate_indices))
                hier = self.compare_hierarchical_indices(self._pad_indices_to_length(query_indices, levels),
                                                         self._pad_indices_to_length(candidate_indices, levels))
            emb = self._calculate_embedding_cosine_similarity(self._extract_original_embedding(query_embedding),
                                                              self._extract_original_embedding(candidate_frame))
            spatial = self._calculate_spatial_locality_similarity(query_embedding, candidate_frame)
            w = self._get_similarity_weights()
            return w["hierarchical"] * hier + w["embedding"] * emb + w["spatial"] * spatial
        return float(comprehensive_scores(np.asarray(candidate_frame, dtype=np.float32)[None], np.asarray(query_embedding, dtype=np.float32),
                                          self._split(query_embedding)[0], device=self._device)[0, 0].item())

    def calculate_embedding_similarity(self, query_embedding: np.ndarray, cached_frames: Dict[int, np.ndarray]) -> List[Tuple[int, float]]:
        """rag/search/engine.py:478-514: comprehensive similarity of every cached frame, sorted descending
        (stable: ties keep the dictionary order)."""
        if query_embedding.size == 0 or not cached_frames:
            return []
        if self.strict:
            query_indices = self._extract_hierarchical_indices(query_embedding)
            sims = [(k, self._calculate_comprehensive_similarity(query_embedding, query_indices, f, k)) for k, f in cached_frames.items()]
            sims.sort(key=lambda x: x[1], reverse=True)
            return sims
        keys = list(cached_frames.keys())
        frames = np.stack([np.asarray(cached_frames[k], dtype=np.float32) for k in keys])
        sc = comprehensive_scores(frames, np.asarray(query_embedding, dtype=np.float32), self._split(query_embedding)[0],
                                  device=self._device)[0].cpu().numpy()
        order = np.argsort(-sc, kind="stable")
        return [(keys[i], float(sc[i])) for i in order]

    def compare_hierarchical_indices(self, query_indices: np.ndarray, candidate_indices: np.ndarray) -> float:
        """rag/search/engine.py:994-1023 (+ multi-level weights :1053-1138)."""
        if query_indices.size == 0 or candidate_indices.size == 0:
            return 0.0
        if query_indices.shape != candidate_indices.shape:
            raise ValueError("Query and candidate indices must have the same shape")
        if query_indices.ndim == 1:
            return self._compare_single_level_indices(query_indices, candidate_indices)
        if query_indices.ndim != 2:
            raise ValueError("Indices must be 1D or 2D arrays")
        L = query_indices.shape[0]
        w = self._calculate_granularity_weights(L)
        tot, tw = 0.0, 0.0
        for l in range(L):
            if query_indices.shape[1] == 0:
                continue
            tot += self._compare_single_level_indices(query_indices[l], candidate_indices[l]) * w[l]
            tw += w[l]
        return tot / tw if tw else 0.0

    @staticmethod
    def _calculate_granularity_weights(num_levels: int) -> np.ndarray:
        if num_levels <= 0:
            return np.array([])
        if num_levels == 1:
            return np.array([1.0])
        w = np.array([8.0 ** (num_levels - i - 1) for i in range(num_levels)])
        w = w / w.sum()
        w[0] *= 2.0
        return w / w.sum()


# ------------------------------------------------------------------------------------------
# core engine surface (core/search_engine.py)
# ------------------------------------------------------------------------------------------
@dataclass
class SearchResult:
    """models.py:40-52"""
    model: object
    similarity_score: float
    matching_indices: Dict[int, float]
    reconstruction_error: float

    def __post_init__(self):
        if self.similarity_score < 0 or self.similarity_score > 1:
            raise ValueError("Similarity score must be between 0 and 1")
        if self.reconstruction_error < 0:
            raise ValueError("Reconstruction error must be non-negative")


class ProgressiveSimilaritySearchEngine:
    """core/search_engine.py:23: per-level similarities run on the device for the whole
    candidate pool at once; the (tiny) filter / sort bookkeeping stays on the host."""

    _result_cls = None             # dropin.install() points this at the reference's models.SearchResult

    def __init__(self, similarity_threshold: float = 0.1, max_candidates_per_level: int = 100, device=None):
        self.similarity_threshold = similarity_threshold
        self.max_candidates_per_level = max_candidates_per_level
        self._device = device

    def _result(self, *fields):
        return (self._result_cls or SearchResult)(*fields)

    def _parse_index_structure(self, indices: np.ndarray, total_space: int):
        return plans.core_levels(len(indices), total_space)

    def _level_sims(self, query_indices: np.ndarray, cand_arrays: Sequence[np.ndarray]) -> np.ndarray:
        """[N, n_query_levels] similarities (levels a candidate lacks score 0.0)."""
        d = dev.require_cuda(self._device)
        q = np.asarray(query_indices, dtype=np.float64)
        q_levels = plans.core_levels(len(q), len(q))
        nl = len(q_levels)
        out = np.zeros((len(cand_arrays), nl))
        if nl == 0 or len(cand_arrays) == 0:
            return out
        tq = torch.from_numpy(q).to(d)
        groups = getattr(cand_arrays, "groups", None)        # frames.QuantizedModelBatch: indices already stacked on the device
        by_len: Dict[int, List[int]] = {}
        if groups is None:
            for i, c in enumerate(cand_arrays):
                by_len.setdefault(len(c), []).append(i)
        else:
            by_len = {S: rows for S, (rows, _) in groups.items()}
        for S, rows in by_len.items():
            if S == 0:
                continue
            c_levels = plans.core_levels(S, S)
            n_cmp = min(nl, len(c_levels))
            if n_cmp == 0:
                continue
            qs = np.array([q_levels[l][1] for l in range(n_cmp)], dtype=np.int32)
            cs = np.array([c_levels[l][1] for l in range(n_cmp)], dtype=np.int32)
            ln = np.array([min(q_levels[l][2] - q_levels[l][1], c_levels[l][2] - c_levels[l][1]) for l in range(n_cmp)],
                          dtype=np.int32)
            if groups is None:
                cand = torch.from_numpy(np.stack([np.asarray(cand_arrays[i], dtype=np.float64) for i in rows])).to(d)
            else:
                cand = groups[S][1].to(d)
            sims = torch.empty((len(rows), n_cmp), dtype=torch.float64, device=d)
            tqs, tcs, tln = (torch.from_numpy(a).to(d) for a in (qs, cs, ln))
            with torch.cuda.device(d):
                check(lib.hq_core_level_sims(dev.ptr(cand), len(rows), S, S, dev.ptr(tq), dev.ptr(tqs), dev.ptr(tcs),
                                             dev.ptr(tln), n_cmp, dev.ptr(sims), dev.stream_ptr()))
            out[np.asarray(rows), :n_cmp] = sims.cpu().numpy()
        return out

    @staticmethod
    def _pool_indices(candidate_pool):
        """the candidates' index vectors: a frames.QuantizedModelBatch keeps them stacked on the device"""
        return candidate_pool if hasattr(candidate_pool, "groups") else [c.hierarchical_indices for c in candidate_pool]

    def compare_indices_at_level(self, query_indices: np.ndarray, candidate_indices: np.ndarray, level: int) -> float:
        if len(query_indices) == 0 or len(candidate_indices) == 0:
            return 0.0
        sims = self._level_sims(query_indices, [candidate_indices])
        return float(sims[0, level]) if level < sims.shape[1] else 0.0

    def _overall(self, sims: np.ndarray) -> np.ndarray:
        L = sims.shape[1]
        w = 1.0 / (np.arange(L) + 1.0)
        return np.clip((sims * w).sum(1) / w.sum(), 0.0, 1.0)

    def brute_force_search(self, query_indices: np.ndarray, candidate_pool: List, max_results: int) -> List[SearchResult]:
        if len(query_indices) == 0 or not candidate_pool:
            return []
        sims = self._level_sims(query_indices, self._pool_indices(candidate_pool))
        if sims.shape[1] == 0:
            overall = np.zeros(len(candidate_pool))
        else:
            overall = self._overall(sims)
        order = np.argsort(-overall, kind="stable")[:max_results]
        return [self._result(candidate_pool[i], float(overall[i]), {l: float(sims[i, l]) for l in range(sims.shape[1])}, 0.0)
                for i in order]

    def progressive_search(self, query_indices: np.ndarray, candidate_pool: List, max_results: int) -> List[SearchResult]:
        if len(query_indices) == 0 or not candidate_pool:
            return []
        sims = self._level_sims(query_indices, self._pool_indices(candidate_pool))
        L = sims.shape[1]
        if L == 0:
            return []
        w = 1.0 / (np.arange(L) + 1.0)
        cur = np.arange(len(candidate_pool))
        for lvl in range(L):
            if len(cur) <= self.max_candidates_per_level:
                break
            ls = sims[cur, lvl]
            combined = (sims[cur, : lvl + 1] * w[: lvl + 1]).sum(1) / w[: lvl + 1].sum()
            keep = ls >= self.similarity_threshold
            kept, kc = cur[keep], combined[keep]
            nxt = kept[np.argsort(-kc, kind="stable")[: self.max_candidates_per_level]]
            if len(nxt) == 0 and len(cur) > 0:
                nxt = cur[[int(np.argmax(ls))]]
            cur = nxt
        overall = self._overall(sims[cur])
        order = np.argsort(-overall, kind="stable")[:max_results]
        res = []
        for j in order:
            i = int(cur[j])
            s = float(overall[j])
            res.append(self._result(candidate_pool[i], s, {l: float(sims[i, l]) for l in range(L)}, max(0.0, 1.0 - s)))
        return res
