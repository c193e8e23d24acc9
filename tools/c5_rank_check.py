"""Diagnostic: time the C5 shard search for the data every rank of an 8-GPU run generates (seed 1234 + rank)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hilbert_quantization_b200 as hq
from hilbert_quantization_b200 import search as S
from bench import make_shard, make_queries_host

dev = torch.device("cuda")
rows, D, Q = 12_500_000, 768, 4096
head = make_shard(torch, 131072, D, 1234, dev)[: Q // 2].cpu().numpy()
q = torch.from_numpy(make_queries_host(head, Q, D)).to(dev)
for r in [int(x) for x in os.environ.get("RANKS", "0,2,5,7").split(",")]:
    emb = make_shard(torch, rows, D, 1234 + r, dev)
    db = hq.EmbeddingDatabase(emb, id_base=r * rows)
    hq.search_batch(db, q, 10)
    torch.cuda.synchronize()
    S.PHASE_TIMER = S.PhaseTimer()
    t0 = time.perf_counter()
    hq.search_batch(db, q, 10)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ph = S.PHASE_TIMER.totals_ms(); S.PHASE_TIMER = None
    print(f"rank-{r} data: wall {wall:.1f} ms, phases {ph}", flush=True)
    del db, emb
    torch.cuda.empty_cache()
