"""Row-sharded search across the GPUs of one box: local progressive top-k per shard, ONE
all-gather of [Q, k] (score, id) pairs over NCCL, then the merge kernel.

Two scopes for the ratio cut of the filter:
  * `filter_scope="shard"` (default, the benchmarked "single all-gather" path): every rank runs
    the reference's filter on its own rows, exactly what the oracle does when given the same
    shard boundaries;
  * `filter_scope="global"`: parity with the reference's single candidate list.  Per level the
    ranks all-reduce the candidate / pass counts, find the exact global cut score with two
    all-reduced histograms over the float32 bit pattern (high then low 16 bits) and resolve
    exact ties by global row id with one small all-gather (`global_ratio_cut`).
Works with the gloo backend on CPU tensors for the host-logic tests (`merge_on_host=True`)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _device as dev


def shard_bounds(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`: [r*N/P, (r+1)*N/P) with the remainder spread over the first ranks."""
    base, rem = divmod(total_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_topk_host(ids: np.ndarray, scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Reference semantics of the merge (score desc, ties -> lower id); ids/scores [P, Q, k]."""
    P, Q, kk = ids.shape
    flat_i = ids.transpose(1, 0, 2).reshape(Q, P * kk)
    flat_s = scores.transpose(1, 0, 2).reshape(Q, P * kk)
    out_i = np.full((Q, k), -1, dtype=np.int64)
    out_s = np.full((Q, k), -1.0, dtype=np.float32)
    for q in range(Q):
        valid = flat_i[q] >= 0
        i, s = flat_i[q][valid], flat_s[q][valid]
        order = np.lexsort((i, -s))[:k]
        out_i[q, : len(order)] = i[order]
        out_s[q, : len(order)] = s[order]
    return out_i, out_s


def allgather_merge(local_ids: torch.Tensor, local_scores: torch.Tensor, k: int, group=None,
                    merge_on_host: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-shard [Q, k] results and merge them on every rank.  Results that come from
    `search.packed_result_buffers` (what `search_batch` returns) travel in ONE all-gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_ids, local_scores
    Q = local_ids.shape[0]
    n = Q * k
    packed = (local_ids.dtype == torch.int64 and local_scores.dtype == torch.float32 and local_ids.is_contiguous()
              and local_scores.is_contiguous() and local_ids.untyped_storage().data_ptr() == local_scores.untyped_storage().data_ptr()
              and local_scores.data_ptr() == local_ids.data_ptr() + 8 * n)
    if packed:
        # search_batch returns ids and scores as two views of one buffer: ONE all-gather ships both
        block = torch.empty(0, dtype=torch.int32, device=local_ids.device).set_(
            local_ids.untyped_storage(), local_ids.storage_offset() * 2, (3 * n,))
        g = torch.empty(world * 3 * n, dtype=torch.int32, device=local_ids.device)
        dist.all_gather_into_tensor(g, block, group=group)
        g2 = g.view(world, 3 * n)
        g_ids = g2[:, : 2 * n]                       # int32 pairs = int64 ids, shard stride 3n int32 = 1.5n int64
        g_sc = g2[:, 2 * n:].view(torch.float32)
        if merge_on_host or not local_ids.is_cuda:
            gi = g_ids.contiguous().view(torch.int64).view(world, Q, k)
            i, s = merge_topk_host(gi.cpu().numpy(), g_sc.contiguous().view(world, Q, k).cpu().numpy(), k)
            return torch.from_numpy(i).to(local_ids.device), torch.from_numpy(s).to(local_scores.device)
        if (3 * n) % 2 == 0:                         # shard blocks stay 8-byte aligned
            from ._lib import check, lib
            from .search import packed_result_buffers
            out_i, out_s = packed_result_buffers(Q, k, local_ids.device)
            with torch.cuda.device(local_ids.device):
                check(lib.hq_topk_merge_strided(g.data_ptr(), g.data_ptr() + 8 * n, world, Q, k, (3 * n) // 2, 3 * n,
                                                dev.ptr(out_i), dev.ptr(out_s), dev.stream_ptr()))
            return out_i, out_s
        g_ids = g_ids.contiguous().view(torch.int64).view(world, Q, k)
        g_sc = g_sc.contiguous().view(world, Q, k)
    else:
        g_ids = torch.empty((world * Q, k), dtype=local_ids.dtype, device=local_ids.device)
        g_sc = torch.empty((world * Q, k), dtype=local_scores.dtype, device=local_scores.device)
        dist.all_gather_into_tensor(g_ids, local_ids.contiguous(), group=group)
        dist.all_gather_into_tensor(g_sc, local_scores.contiguous(), group=group)
        g_ids, g_sc = g_ids.view(world, Q, k), g_sc.view(world, Q, k)
        if merge_on_host or not local_ids.is_cuda:
            i, s = merge_topk_host(g_ids.cpu().numpy(), g_sc.cpu().numpy(), k)
            return torch.from_numpy(i).to(local_ids.device), torch.from_numpy(s).to(local_scores.device)
    from ._lib import check, lib
    out_i = torch.empty((Q, k), dtype=torch.int64, device=local_ids.device)
    out_s = torch.empty((Q, k), dtype=torch.float32, device=local_ids.device)
    with torch.cuda.device(local_ids.device):
        check(lib.hq_topk_merge(dev.ptr(g_ids), dev.ptr(g_sc), world, Q, k, dev.ptr(out_i), dev.ptr(out_s), dev.stream_ptr()))
    return out_i, out_s


class MergePipeline:
    """The all-gather + merge of batch i on a communication stream while the main stream already searches batch i + 1.

    A step of the row-sharded search ends with ~0.1 ms of NCCL latency and a small merge kernel; with shards of 125 K
    rows (1 M rows over 8 GPUs) the search itself takes 0.9 ms, so a serial step spends a tenth of its time waiting for
    120 KB of results.  `submit` orders the collective after the search (event on the main stream), runs it on the
    pipeline's stream and returns the merged (ids, scores) device tensors, valid once `done` (the returned event) has
    completed or after `drain()`.  Every rank must submit the same sequence of batches."""

    def __init__(self, device, group=None, depth: int = 2):
        self.device, self.group, self.depth = torch.device(device), group, max(1, int(depth))
        self.stream = torch.cuda.Stream(device=self.device)
        self._inflight: list = []

    def reserve(self):
        """The main stream waits until fewer than `depth` merges are in flight.  Call it BEFORE producing the next batch's
        results when those live in reused buffers (SearchGraph output buffers: two graphs alternate, depth = 2)."""
        main = torch.cuda.current_stream(self.device)
        while len(self._inflight) >= self.depth:
            main.wait_event(self._inflight.pop(0))

    def submit(self, local_ids: torch.Tensor, local_scores: torch.Tensor, k: int):
        main = torch.cuda.current_stream(self.device)
        self.reserve()                               # bound the work queued on the communication stream
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            out_i, out_s = allgather_merge(local_ids, local_scores, k, self.group)
            for t in (local_ids, local_scores, out_i, out_s):
                t.record_stream(self.stream)         # the caching allocator must not hand these to the main stream early
            done = torch.cuda.Event()
            done.record(self.stream)
        self._inflight.append(done)
        return out_i, out_s, done

    def drain(self):
        """The main stream waits for every submitted merge."""
        main = torch.cuda.current_stream(self.device)
        for ev in self._inflight:
            main.wait_event(ev)
        self._inflight.clear()


def _all_reduce(t: torch.Tensor, op, group) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def _all_gather_padded(t: torch.Tensor, fill, group) -> torch.Tensor:
    """[Q, m_local] -> [Q, world * m_max] (ranks padded to the longest with `fill`)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return t
    world = dist.get_world_size(group)
    m = torch.tensor([t.shape[1]], dtype=torch.int64, device=t.device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    m_max = int(m.item())
    pad = torch.full((t.shape[0], m_max), fill, dtype=t.dtype, device=t.device)
    pad[:, : t.shape[1]] = t
    out = torch.empty((world * t.shape[0], m_max), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out.view(world, t.shape[0], m_max).permute(1, 0, 2).reshape(t.shape[0], world * m_max)


def global_ratio_cut(scores: torch.Tensor, passed: torch.Tensor, n_alive_local: torch.Tensor, ratio: float,
                     id_base: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's ratio cut over the GLOBAL candidate list of each query (rag/search/engine.py:272-287).

    scores [Q, N_local] float32 (>= 0), passed [Q, N_local] bool (alive and score >= threshold),
    n_alive_local [Q] rows of this shard that entered the level.  Returns (keep [Q, N_local] bool,
    n_out [Q] int64 global survivors).  keep = passed rows among the cap best of the global list
    (score desc, ties -> lower global row id), cap = max(1, int(n_alive_global * ratio)).
    Pure tensor code + three small collectives per level; runs on CPU tensors (gloo) or CUDA (NCCL)."""
    Q, N = scores.shape
    dev_ = scores.device
    n_alive = _all_reduce(n_alive_local.to(torch.int64).clone(), dist.ReduceOp.SUM, group)
    n_pass = _all_reduce(passed.sum(1).to(torch.int64), dist.ReduceOp.SUM, group)
    cap = torch.clamp((n_alive.to(torch.float64) * ratio).floor().to(torch.int64), min=1)
    need = n_pass > cap                                           # identical on every rank
    keep = passed.clone()
    n_out = torch.minimum(n_pass, cap)
    if not bool(need.any()):
        return keep, n_out
    rows_q = torch.nonzero(need).flatten()
    key = scores[rows_q].contiguous().view(torch.int32).to(torch.int64)       # non-negative floats: bit order == value order
    ok = passed[rows_q]
    capq = cap[rows_q]
    nq = rows_q.numel()

    def cut_digit(digit, valid, want, bins):
        """largest digit d with #(valid, digit >= d) >= want; returns (d, count strictly above d)"""
        hist = torch.zeros((nq, bins), dtype=torch.int32, device=dev_)
        hist.scatter_add_(1, torch.where(valid, digit, torch.zeros_like(digit)), valid.to(torch.int32))
        _all_reduce(hist, dist.ReduceOp.SUM, group)
        from_top = hist.to(torch.int64).flip(1).cumsum(1).flip(1)                                 # #(digit >= d)
        d = (from_top >= want[:, None]).to(torch.int64).sum(1) - 1                 # from_top is non-increasing in d
        above = from_top.gather(1, torch.clamp(d + 1, max=bins - 1)[:, None])[:, 0]
        above = torch.where(d + 1 < bins, above, torch.zeros_like(above))
        return d, above

    hi, lo = key >> 16, key & 0xffff
    d1, above1 = cut_digit(hi, ok, capq, 1 << 15)
    in1 = ok & (hi == d1[:, None])
    d2, above2 = cut_digit(lo, in1, capq - above1, 1 << 16)
    k_star = (d1 << 16) | d2
    r_ties = capq - above1 - above2                                                 # ties to keep (>= 1), lowest global ids
    tie = ok & (key == k_star[:, None])
    gid = torch.arange(N, device=dev_, dtype=torch.int64)[None, :] + id_base
    big = torch.iinfo(torch.int64).max
    m_loc = int(tie.sum(1).max().item()) if nq else 0
    tie_ids = torch.where(tie, gid.expand(nq, N), torch.full((1, 1), big, dtype=torch.int64, device=dev_)).sort(1).values[:, :max(m_loc, 1)]
    all_ids = _all_gather_padded(tie_ids.contiguous(), big, group).sort(1).values
    g_star = all_ids.gather(1, torch.clamp(r_ties - 1, min=0, max=all_ids.shape[1] - 1)[:, None])[:, 0]
    keep_q = ok & ((key > k_star[:, None]) | (tie & (gid <= g_star[:, None])))
    keep[rows_q] = keep_q
    return keep, n_out


def global_ratio_cut_device(scores: torch.Tensor, mask: torch.Tensor, N: int, n_alive_local: torch.Tensor,
                            n_pass_local: torch.Tensor, ratio: float, id_base: int, group=None) -> torch.Tensor:
    """`global_ratio_cut` on the device without [Q, N] temporaries: `scores` [Q, >= N] float32 are the level scores of
    hq_filter_level, `mask` [Q, words] int32 its pass bits, rewritten IN PLACE with the survivors of the global cut.
    Per level: two all-reduced [Q, 65536] histograms (high / low 16 bits of the score's bit pattern) locate the exact cut
    score, one all-gather of the per-shard tie counts shares out the ties by ascending global row id (row shards are
    contiguous id ranges, so a shard's quota is what the shards with smaller ids leave).  Returns n_out [Q] int64.
    Every rank must call it, an empty shard (N == 0) included."""
    import ctypes as C
    from ._lib import check, lib
    Q = int(n_alive_local.shape[0])
    d = n_alive_local.device
    n_alive = _all_reduce(n_alive_local.to(torch.int64).clone(), dist.ReduceOp.SUM, group)
    n_pass = _all_reduce(n_pass_local.to(torch.int64).clone(), dist.ReduceOp.SUM, group)
    cap = torch.clamp((n_alive.to(torch.float64) * ratio).floor().to(torch.int64), min=1)
    need = n_pass > cap                                           # identical on every rank
    n_out = torch.minimum(n_pass, cap)
    if Q == 0 or not bool(need.any()):
        return n_out
    need32 = need.to(torch.int32).contiguous()
    st = dev.stream_ptr()
    s_ptr, s_stride = (dev.ptr(scores), scores.stride(0)) if N > 0 else (None, 0)
    m_ptr, m_stride = (dev.ptr(mask), mask.stride(0)) if N > 0 else (None, 0)
    hist = torch.zeros((Q, 65536), dtype=torch.int32, device=d)
    d1, above1, d2, above2 = (torch.zeros(Q, dtype=torch.int64, device=d) for _ in range(4))
    with torch.cuda.device(d):
        if N > 0:
            check(lib.hq_gcut_hist(s_ptr, s_stride, N, Q, m_ptr, m_stride, dev.ptr(need32), None, dev.ptr(hist), st))
        _all_reduce(hist, dist.ReduceOp.SUM, group)
        check(lib.hq_gcut_scan(dev.ptr(hist), Q, dev.ptr(need32), dev.ptr(cap), dev.ptr(d1), dev.ptr(above1), st))
        d1_32 = d1.to(torch.int32).contiguous()
        hist.zero_()
        if N > 0:
            check(lib.hq_gcut_hist(s_ptr, s_stride, N, Q, m_ptr, m_stride, dev.ptr(need32), dev.ptr(d1_32), dev.ptr(hist), st))
        _all_reduce(hist, dist.ReduceOp.SUM, group)
        want2 = (cap - above1).contiguous()
        check(lib.hq_gcut_scan(dev.ptr(hist), Q, dev.ptr(need32), dev.ptr(want2), dev.ptr(d2), dev.ptr(above2), st))
        k_star = ((d1 << 16) | d2).contiguous()
        r_ties = cap - above1 - above2                            # ties to keep over all shards (>= 1)
        ties = torch.zeros(Q, dtype=torch.int64, device=d)
        if N > 0:
            check(lib.hq_gcut_ties(s_ptr, s_stride, N, Q, m_ptr, m_stride, dev.ptr(need32), dev.ptr(k_star), dev.ptr(ties), st))
        before = torch.zeros(Q, dtype=torch.int64, device=d)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            world = dist.get_world_size(group)
            mine = torch.cat([torch.tensor([int(id_base)], dtype=torch.int64, device=d), ties])
            allr = torch.empty((world, Q + 1), dtype=torch.int64, device=d)
            dist.all_gather_into_tensor(allr.view(-1), mine.contiguous(), group=group)
            lower = (allr[:, 0] < int(id_base)) | ((allr[:, 0] == int(id_base)) & (torch.arange(world, device=d) < dist.get_rank(group)))
            before = (allr[:, 1:] * lower[:, None].to(torch.int64)).sum(0)
        quota = torch.minimum(torch.clamp(r_ties - before, min=0), ties).contiguous()
        if N > 0:
            check(lib.hq_gcut_apply(s_ptr, s_stride, N, Q, m_ptr, m_stride, dev.ptr(need32), dev.ptr(k_star), dev.ptr(quota), st))
    return n_out


class ShardedSearch:
    """One instance per rank.  `local_embeddings` are this rank's rows [start, end) of the global database."""

    def __init__(self, local_embeddings, global_row_start: int, n: Optional[int] = None, device=None, group=None):
        from .search import EmbeddingDatabase
        self.group = group
        self.db = EmbeddingDatabase(local_embeddings, n=n, device=device, id_base=global_row_start)

    def search(self, queries, k: int = 10, filter_scope: str = "shard", **kw):
        from .search import search_batch
        ids, scores = search_batch(self.db, queries, k, filter_scope=filter_scope, group=self.group, **kw)
        return allgather_merge(ids, scores, k, self.group)

    def search_pipelined(self, query_batches, k: int = 10, **kw):
        """Throughput path for device-resident query batches: yields (ids, scores, done_event) per batch, the merge of
        batch i running on a communication stream under the search of batch i + 1 (see `MergePipeline`).  The caller waits
        for `done_event` (or synchronises the device) before reading a batch's result."""
        from .search import search_batch
        pipe = MergePipeline(self.db.device, self.group)
        for q in query_batches:
            ids, scores = search_batch(self.db, q, k, group=self.group, **kw)
            yield pipe.submit(ids, scores, k)
        pipe.drain()

    def search_stream(self, host_batches, k: int = 10, *, depth: int = 2, **kw):
        """`search.search_stream` over this rank's shard with the all-gather merge as its post step: every rank feeds the same
        host batches and gets the merged global top-k of each batch back in pinned host memory, copies overlapped with the
        search (the bench's e2e figure at N > 1 is exactly this composition)."""
        from .search import search_stream
        if getattr(self, "_post_stream", None) is None:
            self._post_stream = torch.cuda.Stream(device=self.db.device)
        return search_stream(self.db, host_batches, k, depth=depth, post_stream=self._post_stream,
                             post=lambda ids, scores: allgather_merge(ids, scores, k, self.group), **kw)
