"""Empirical fallback rate of the window-mode filter: how often does a query's ratio cut fall outside its predicted window?
Run with HQ_FILTER_WINDOW_Z=z (window half width in standard deviations of the sample rank).

    HQ_FILTER_WINDOW_Z=4 python tools/window_fallback_rate.py [rows] [dim] [batches]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hilbert_quantization_b200 as hq
from hilbert_quantization_b200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
Q = 1024
g = torch.Generator(device="cuda").manual_seed(1234)
db = torch.randn((N, D), device="cuda", generator=g)
db /= db.norm(dim=1, keepdim=True)
d = hq.EmbeddingDatabase(db)
off = int(lib.hq_filter_fast_fallback_offset(N, Q, C.byref(d.layout)))
total = fb = 0
for b in range(B):
    qs = torch.randn((Q, D), device="cuda", generator=g)
    if b % 2:                                              # half of the batches: queries near database rows
        rows = torch.randint(0, N, (Q,), device="cuda", generator=g)
        qs = db[rows] + 0.1 * qs / D ** 0.5
    hq.search_batch(d, qs, 10, filter_impl="fast")
    torch.cuda.synchronize()
    fb += int(d._filter_scratch[off: off + 4 * Q].view(torch.int32).sum().item())
    total += Q
print(f"z={os.environ.get('HQ_FILTER_WINDOW_Z', '5')} rows={N} dim={D}: {fb} fallbacks in {total} queries ({2 * total} cut checks)")
