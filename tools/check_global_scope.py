"""2+ GPU check of filter_scope="global" (torchrun): the row-sharded search with the global ratio cut must
return exactly what one GPU returns on the whole database (the reference's single candidate list).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_global_scope.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import hilbert_quantization_b200 as hq
from hilbert_quantization_b200.distributed import ShardedSearch, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (N, D, Q, positive) in [(20000, 1536, 16, False), (9001, 768, 10, True)]:
    rng = np.random.default_rng(N)
    if positive:
        db = (rng.random((N, D)) + 0.25).astype(np.float32)
        qs = (rng.random((Q, D)) + 0.25).astype(np.float32)
    else:
        db = rng.standard_normal((N, D)).astype(np.float32)
        qs = rng.standard_normal((Q, D)).astype(np.float32)
    db[N // 2 + 3] = db[5]
    qs[0] = db[5]
    lo, hi = shard_bounds(N, world, rank)
    sh = ShardedSearch(db[lo:hi], lo)
    ids_g, sc_g = sh.search(qs, 10, filter_scope="global")
    ids_s, sc_s = sh.search(qs, 10, filter_scope="shard")
    full = hq.EmbeddingDatabase(db)
    ids_1, sc_1 = hq.search_batch(full, qs, 10, filter_impl="exact")
    same = torch.equal(ids_g, ids_1) and torch.equal(sc_g, sc_1)
    differs = not torch.equal(ids_s, ids_1)
    if rank == 0:
        print(f"N={N} D={D} Q={Q} positive={positive}: global == single list: {same}; per-shard cut differs from single list: {differs}")
    ok = ok and same
dist.destroy_process_group()
sys.exit(0 if ok else 1)
