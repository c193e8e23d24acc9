"""C1 (25 K x 1536, one query) latency breakdown: device phases (CUDA events) vs host wall time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hilbert_quantization_b200 as hq
from hilbert_quantization_b200 import search as S
from hilbert_quantization_b200._lib import lib
from bench import make_shard, make_queries_host

dev = torch.device("cuda")
emb = make_shard(torch, 25_000, 1536, 1234, dev)
db = hq.EmbeddingDatabase(emb)
q = torch.from_numpy(make_queries_host(emb[:1].cpu().numpy(), 2, 1536)[:1]).to(dev)
for _ in range(20):
    hq.search_batch(db, q, 10)
torch.cuda.synchronize()
lib.hq_launch_count(1)
S.PHASE_TIMER = S.PhaseTimer()
t0 = time.perf_counter()
for _ in range(50):
    ids, sc = hq.search_batch(db, q, 10)
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 50 * 1e3
ph = S.PHASE_TIMER.totals_ms()
S.PHASE_TIMER = None
print("wall per call (sync each)", round(wall, 4), "ms; device phases per call:", {k: round(v / 50, 4) for k, v in ph.items()},
      "launches per call", lib.hq_launch_count(1) / 50)
t0 = time.perf_counter()
for _ in range(200):
    ids, sc = hq.search_batch(db, q, 10)
torch.cuda.synchronize()
print("wall per call (no sync between calls)", round((time.perf_counter() - t0) / 200 * 1e3, 4), "ms")
