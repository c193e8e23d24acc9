import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hilbert_quantization_b200 as hq
from bench import make_shard
dev = torch.device("cuda")
D = 768
for rows, seed, base in [(2_000_000, 1234, 0), (2_000_000, 1236, 0), (2_000_000, 1236, 25_000_000), (12_500_000, 1236, 0)]:
    emb = make_shard(torch, rows, D, seed, dev)
    db = hq.EmbeddingDatabase(emb, id_base=base)
    lens = db.lens.cpu().numpy().astype(np.int64)
    keff = [int(db.layout.lvl_keff[l]) for l in range(db.num_levels)]
    bad = np.nonzero((lens != np.array(keff)[None, :]).any(axis=1))[0]
    print(rows, seed, base, "fast_filter_ok", db.fast_filter_ok, "tc", db.tc_packed is not None, "bad rows", len(bad), bad[:5], flush=True)
    if len(bad):
        r = int(bad[0])
        print("   lens", lens[r], "keff", keff, "idx row", db.idx[r].cpu().numpy(), "norm", float(emb[r].norm()), "emb has nan", bool(torch.isnan(emb[r]).any()))
    del db, emb
    torch.cuda.empty_cache()
