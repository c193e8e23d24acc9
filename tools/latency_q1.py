"""Launch list of ONE single-query search on the 1 M x 1536 database (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hilbert_quantization_b200 as hq

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1234)
emb = torch.randn((rows, 1536), device=dev, generator=g)
emb /= emb.norm(dim=1, keepdim=True)
db = hq.EmbeddingDatabase(emb, device=dev)
q = emb[:1] + 0.01 * torch.randn_like(emb[:1])
for _ in range(3):
    ids, sc = hq.search_batch(db, q, 10)
torch.cuda.synchronize()
print(ids, sc)
