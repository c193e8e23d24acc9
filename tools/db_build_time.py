"""Time EmbeddingDatabase construction (index rows, norms, bf16 unit rows, packed filter operand) for one shard."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hilbert_quantization_b200 as hq
from bench import make_shard

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
emb = make_shard(torch, rows, dim, 1234, torch.device("cuda"))
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    db = hq.EmbeddingDatabase(emb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gb = rows * dim * 4 / 1e9
    print(f"build {rows} x {dim}: {dt * 1e3:.2f} ms ({gb / dt:.0f} GB/s of fp32 embeddings)")
    del db
