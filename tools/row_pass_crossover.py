"""Filter time of small batches with the CUDA-core row pass (csrc/hq_filter_rows.cu) and with the tensor-core window pass:
    python tools/row_pass_crossover.py [rows]      (one line per batch size)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hilbert_quantization_b200 as hq
from hilbert_quantization_b200.search import prepare_queries, progressive_filter_fast

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1234)
emb = torch.randn((rows, 1536), device=dev, generator=g)
emb /= emb.norm(dim=1, keepdim=True)
db = hq.EmbeddingDatabase(emb, device=dev)
words = (rows + 31) // 32
for Q in (1, 2, 4, 8):
    qs = emb[:Q] + 0.01 * torch.randn((Q, 1536), device=dev, generator=g)
    q, q_idx, q_lens, q_norms = prepare_queries(db, qs)
    mask = torch.zeros((Q, (words + 7) // 8 * 8), dtype=torch.int32, device=dev)[:, :words]
    out = {}
    for row_pass in (True, False):
        for _ in range(3):
            progressive_filter_fast(db, q_idx, mask, row_pass=row_pass)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            progressive_filter_fast(db, q_idx, mask, row_pass=row_pass)
        e1.record()
        torch.cuda.synchronize()
        out[row_pass] = e0.elapsed_time(e1) / 20
    print(f"Q={Q}: filter {out[True]*1e3:.0f} us with the row pass, {out[False]*1e3:.0f} us with the tensor-core window pass")
