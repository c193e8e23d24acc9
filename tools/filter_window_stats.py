"""Window-mode diagnostics of the fast filter on one GPU: list sizes, window widths, counters and fallbacks per query.

    python tools/filter_window_stats.py [rows] [dim] [queries]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hilbert_quantization_b200 as hq
from hilbert_quantization_b200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
g = torch.Generator(device="cuda").manual_seed(1234)
db = torch.randn((N, D), device="cuda", generator=g)
db /= db.norm(dim=1, keepdim=True)
qs = torch.randn((Q, D), device="cuda", generator=g)
qs[: Q // 2] = db[: Q // 2] + 0.1 * torch.randn((Q // 2, D), device="cuda", generator=g) / D ** 0.5 * D ** 0.5 * 0.026
d = hq.EmbeddingDatabase(db)
print("mode", int(lib.hq_filter_fast_mode(N, Q, C.byref(d.layout))), "levels", int(d.layout.L))
budget = int(sys.argv[4]) if len(sys.argv) > 4 else (24 << 30)
while Q > 128 and int(lib.hq_filter_fast_scratch_bytes(N, Q, C.byref(d.layout))) > budget:      # one chunk of search_batch
    Q = max(128, (Q // 2 + 127) // 128 * 128)
qs = qs[:Q].contiguous()
print("queries in one filter chunk:", Q)
ids, sc, mask = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="fast", filter_scratch_bytes=budget)
torch.cuda.synchronize()
out = (C.c_int64 * 10)()
rc = lib.hq_filter_fast_window_layout(N, Q, C.byref(d.layout), out)
assert rc == 0
s = d._filter_scratch
off = list(out)
n_segs, seg_cap, stride = off[7], off[8], off[9]
win = s[off[1]: off[1] + 16 * Q].view(torch.float32).reshape(4, Q).cpu().numpy()
cnt = s[off[2]: off[2] + 16 * Q].view(torch.int32).reshape(4, Q).cpu().numpy()
c0s = s[off[3]: off[3] + 4 * Q].view(torch.int32).cpu().numpy()
pflag = s[off[4]: off[4] + 4 * Q].view(torch.int32).cpu().numpy()
seg_n = s[off[6]: off[6] + 4 * Q * n_segs].view(torch.int32).reshape(Q, n_segs).cpu().numpy()
fb_off = int(lib.hq_filter_fast_fallback_offset(N, Q, C.byref(d.layout)))
fb = s[fb_off: fb_off + 4 * Q].view(torch.int32).cpu().numpy()
ents = seg_n.sum(1)
print(f"sample stride {stride}, segments {n_segs} x cap {seg_cap}; fallbacks {int(fb.sum())} of {Q}, prediction flags {int(pflag.sum())}")
print("list entries per query: mean %.0f  p50 %.0f  max %d  (%.2f %% of the rows); largest segment %d" %
      (ents.mean(), np.median(ents), ents.max(), 100.0 * ents.mean() / N, seg_n.max()))
print("counters per query (mean): c0 %.0f  above window 1 %.0f  of those k2>=lo2 %.0f  alive for sure %.0f;  sample c0 %.0f" %
      tuple(cnt.mean(1).tolist() + [c0s.mean()]))
print("survivors per query (mask popcount): mean %.0f" % float(hq.search.unpack_mask(mask[:16], N).sum(1).mean()))
for name, a in (("lo1", win[0]), ("hi1", win[1]), ("lo2", win[2]), ("hi2", win[3])):
    fin = np.isfinite(a)
    print(f"{name}: finite {int(fin.sum())} of {Q}, mean {a[fin].mean() if fin.any() else float('nan'):.5f}")
