"""Row-sharded search across the GPUs of one box: local progressive top-k per shard, ONE
all-gather of [Q, k] (score, id) pairs over NCCL, then the merge kernel.

The ratio cut of the filter is applied per shard (`filter_scope="shard"`): every rank runs
the reference's filter on its own rows, exactly what the oracle does when given the same
shard boundaries.  Works with the gloo backend on CPU tensors for the host-logic tests
(`merge_on_host=True`)."""
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
    """All-gather the per-shard [Q, k] results and merge them on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_ids, local_scores
    Q = local_ids.shape[0]
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


class ShardedSearch:
    """One instance per rank.  `local_embeddings` are this rank's rows [start, end) of the global database."""

    def __init__(self, local_embeddings, global_row_start: int, n: Optional[int] = None, device=None, group=None):
        from .search import EmbeddingDatabase
        self.group = group
        self.db = EmbeddingDatabase(local_embeddings, n=n, device=device, id_base=global_row_start)

    def search(self, queries, k: int = 10, **kw):
        from .search import search_batch
        ids, scores = search_batch(self.db, queries, k, **kw)
        return allgather_merge(ids, scores, k, self.group)
