"""Cross-rank parity of the row-sharded search over NCCL (run under torchrun, >= 2 ranks, one GPU each).

  1. filter_scope="shard": merged ids == the oracle run PER SHARD (same shard bounds) and merged on the host
     (rag/search/engine.py:51-95 + :622-660 + :512 per shard, ties -> lower global id);
  2. filter_scope="global": merged ids == the oracle over the SINGLE global candidate list
     (rag/search/engine.py:272-287), with uneven shards and several query chunks per rank (the per-level
     collectives must line up across ranks);
  3. the same with one EMPTY shard;
  4. bf16-only shards (EmbeddingDatabase.from_chunks): sharded result == one rank holding every row.
Rank 0 prints one summary line per check; any mismatch raises (non-zero exit)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hilbert_quantization_b200 as hq                                   # noqa: E402
from hilbert_quantization_b200.distributed import allgather_merge, merge_topk_host, shard_bounds   # noqa: E402
from oracle import hilbert_oracle as O                                   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(2026)
    N, D, Q, k = 60001, 768, 48, 10                                       # 60001: shards of unequal size
    db = rng.standard_normal((N, D)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    db[40000] = db[17]                                                    # an exact tie across shards
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    qs[: Q // 2] = db[: Q // 2] + 0.05 * rng.standard_normal((Q // 2, D)).astype(np.float32)
    qs[1] = db[17]
    n = hq.rag_optimal_dimensions(D)[0]
    lo, hi = shard_bounds(N, world, rank)
    shard = hq.ShardedSearch(db[lo:hi], global_row_start=lo, n=n, device=dev)
    checked = list(range(0, Q, 6))

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    # ---- 1. shard scope vs the oracle per shard ----
    ids, sc = shard.search(qs, k)
    ids_h, sc_h = ids.cpu().numpy(), sc.cpu().numpy()
    for j in checked:
        parts_i, parts_s = [], []
        for r in range(world):
            a, b = shard_bounds(N, world, r)
            iw, sw = O.progressive_search(qs[j], db[a:b], n, k)
            pi = np.full(k, -1, dtype=np.int64); ps = np.full(k, -1.0, dtype=np.float32)
            pi[: len(iw)] = iw + a; ps[: len(iw)] = sw
            parts_i.append(pi[None]); parts_s.append(ps[None])
        wi, ws = merge_topk_host(np.stack(parts_i), np.stack(parts_s), k)
        m = int((wi[0] >= 0).sum())
        assert list(ids_h[j, :m]) == list(wi[0, :m]), (rank, j, ids_h[j], wi[0])
        assert (ids_h[j, m:] == -1).all()
        assert np.abs(sc_h[j, :m] - ws[0, :m]).max() < 5e-7
    assert ids_h[1, 0] == 17 and ids_h[1, 1] == 40000                      # the tie resolves to the lower GLOBAL id
    say(f"shard scope: {len(checked)} queries equal the per-shard oracle merged on the host (world {world})")

    # ---- 2. global scope vs the single-list oracle; 7 queries per chunk -> several rounds of collectives per level ----
    small = 4 * (N // world + 1) * 7                                       # the same on every rank: 7 queries per chunk
    gi, gs = shard.search(qs, k, filter_scope="global", work_bytes=small, rerank="f32")
    gi_h, gs_h = gi.cpu().numpy(), gs.cpu().numpy()
    for j in checked:
        iw, sw = O.progressive_search(qs[j], db, n, k)
        m = len(iw)
        assert list(gi_h[j, :m]) == list(iw), (rank, j, gi_h[j], iw)
        assert np.abs(gs_h[j, :m] - sw).max() < 5e-7
    gi2, gs2 = shard.search(qs, k, filter_scope="global")                  # default chunking, tensor-core rerank
    assert torch.equal(gi2, gi) and (gs2 - gs).abs().max().item() <= 1e-6
    say(f"global scope: {len(checked)} queries equal the single-list oracle (uneven shards, {-(-Q // 7)} query chunks)")

    # ---- 3. one empty shard ----
    a, b = (0, N) if rank == 0 else (N, N)
    lone = hq.ShardedSearch(db[a:b], global_row_start=a, n=n, device=dev)
    ei, es = lone.search(qs, k, filter_scope="global", work_bytes=small, rerank="f32")
    assert torch.equal(ei.cpu(), gi.cpu()) and (es.cpu() - gs.cpu()).abs().max().item() <= 1e-6
    ei, es = lone.search(qs, k)
    one_i, one_s = hq.search_batch(hq.EmbeddingDatabase(db, n=n, device=dev), qs, k)
    assert torch.equal(ei, one_i) and torch.equal(es, one_s)
    say("empty shard: every rank took part in the collectives, results equal the single-shard search")

    # ---- 4. bf16-only shards ----
    lean = hq.EmbeddingDatabase.from_chunks(iter([torch.from_numpy(db[s:min(hi, s + 9000)]).to(dev) for s in range(lo, hi, 9000)]),
                                            hi - lo, D, n=n, device=dev, id_base=lo)
    li, ls = hq.search_batch(lean, qs, k)
    li, ls = allgather_merge(li, ls, k)
    whole = hq.EmbeddingDatabase.from_chunks(iter([torch.from_numpy(db).to(dev)]), N, D, n=n, device=dev)
    wi_, ws_ = hq.search_batch(whole, qs, k, use_filter=False)
    li2, ls2 = hq.search_batch(lean, qs, k, use_filter=False)
    li2, ls2 = allgather_merge(li2, ls2, k)
    assert torch.equal(li2, wi_) and torch.equal(ls2, ws_)                 # no filter: sharding cannot change the result
    assert (li >= -1).all()
    say("bf16-only shards: merged top-k of the shards == one rank holding all rows (unfiltered); filtered search ran")
    dist.barrier()
    dist.destroy_process_group()
    say("nccl_check ok")


if __name__ == "__main__":
    main()
