"""Filter statistics and per-kernel timing for a synthetic shard (diagnostics): pass rates per level and the time of
the fast filter with / without the tensor-core pass."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hilbert_quantization_b200 as hq
from hilbert_quantization_b200.search import FilterTrace
from bench import make_shard, make_queries_host

N, D, Q = int(os.environ.get("N", 2_000_000)), int(os.environ.get("D", 768)), int(os.environ.get("Q", 512))
dev = torch.device("cuda")
emb = make_shard(torch, N, D, 1234, dev)
db = hq.EmbeddingDatabase(emb)
q = torch.from_numpy(make_queries_host(emb[: Q // 2].cpu().numpy(), Q, D)).to(dev)
tr = FilterTrace([], [], [])
hq.search_batch(db, q, 10, trace=tr)
for l in range(len(tr.n_out)):
    a, p, o = tr.n_alive[l].float(), tr.n_pass[l].float(), tr.n_out[l].float()
    print(f"level {l}: alive {a.mean():.0f}  pass {p.mean():.0f} ({(p / a).mean():.3f}, max {(p / a).max():.3f})  out {o.mean():.0f}  cut binds for {(p > o).float().mean():.2f} of queries")
for impl in ("fast", "fast_fp32"):
    for _ in range(2):
        hq.search_batch(db, q, 10, filter_impl=impl)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        hq.search_batch(db, q, 10, filter_impl=impl)
    torch.cuda.synchronize()
    print(impl, f"{(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per batch")
