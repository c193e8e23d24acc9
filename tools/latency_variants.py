"""Single-query latency of the search through SearchGraph for the filter / rerank variants (1 M x 1536 by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hilbert_quantization_b200 as hq
from bench import make_shard

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda")
emb = make_shard(torch, rows, 1536, 1234, dev)
db = hq.EmbeddingDatabase(emb, device=dev)
q = emb[:Q] + 0.01 * torch.randn_like(emb[:Q])
ref = None
for name, kw in (("auto", {}), ("fast_fp32", {"filter_impl": "fast_fp32"}), ("bf16 rerank", {"rerank": "bf16"}),
                 ("fast_fp32 + bf16", {"filter_impl": "fast_fp32", "rerank": "bf16"})):
    try:
        sg = hq.SearchGraph(db, Q, 10, **kw)
        for _ in range(5):
            sg.search(q)
        ids = sg.search(q)[0].clone()
        if ref is None:
            ref = ids
        ms = []
        for _ in range(50):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sg.search(q); e1.record(); e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        print(f"{name:20s} p50 {np.median(ms):.3f} ms  p99 {np.percentile(ms, 99):.3f}  same ids {bool(torch.equal(ids, ref))}")
    except Exception as ex:
        print(name, "failed:", str(ex)[:200])
