"""Time the fused map(+index) pass in its variants (B200): map only, map + variant-C index, index only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hilbert_quantization_b200 as hq
from hilbert_quantization_b200.index import fused_pass, plans

N, D, n = 262144, int(os.environ.get("D", "1536")), int(os.environ.get("NSIDE", "64"))
dev = torch.device("cuda")
emb = torch.randn((N, D), device=dev)
grids = torch.empty((N, n * n), device=dev)
plan, widths, ml = plans.c_plan(n, "compact")
idx = torch.empty((N, len(plan)), device=dev)


def timeit(f, reps=8):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = {
    "map only": (lambda: fused_pass(emb, 0, n, D, plan=None, plan_key=None, min_level=99, grid_out=grids, idx_out=None), 4 * D + 4 * n * n),
    "map + index C": (lambda: fused_pass(emb, 0, n, D, plan=plan, plan_key=("C", n, "compact"), min_level=ml, grid_out=grids, idx_out=idx),
                      4 * D + 4 * n * n + 4 * len(plan)),
    "index C only": (lambda: fused_pass(emb, 0, n, D, plan=plan, plan_key=("C", n, "compact"), min_level=ml, grid_out=None, idx_out=idx),
                     4 * D + 4 * len(plan)),
}
for name, (f, bpi) in cases.items():
    try:
        ms = timeit(f)
        print(f"{name:16s} {ms:8.4f} ms  {N * bpi / ms / 1e6:8.0f} GB/s  ({bpi} B/item)")
    except Exception as e:
        print(name, "failed:", e)
