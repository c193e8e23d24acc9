"""Where does the host-clock e2e time go?  Per-batch host timestamps of hq.search_stream against the device-resident loop."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hilbert_quantization_b200 as hq

N, D, Q, k = int(os.environ.get("ROWS", 1_000_000)), 1536, 1024, 10
g = torch.Generator(device="cuda").manual_seed(1)
emb = torch.randn((N, D), generator=g, device="cuda"); emb /= emb.norm(dim=1, keepdim=True)
db = hq.EmbeddingDatabase(emb)
q = torch.randn((Q, D), generator=g, device="cuda"); q /= q.norm(dim=1, keepdim=True)
qp = q.cpu().pin_memory()
for _ in range(5):
    hq.search_batch(db, q, k)
torch.cuda.synchronize()
for steps in (20, 20, 100):
    t0 = time.perf_counter()
    for _ in range(steps):
        hq.search_batch(db, q, k)
    torch.cuda.synchronize()
    dev_ms = (time.perf_counter() - t0) * 1e3
    stamps = []
    t0 = time.perf_counter()
    for ids, sc in hq.search_stream(db, (qp for _ in range(steps)), k):
        stamps.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    d = np.diff([0.0] + stamps) * 1e3
    print(f"steps {steps}: device loop {dev_ms:.1f} ms ({dev_ms/steps:.2f}/step)  stream {e2e_ms:.1f} ms ({e2e_ms/steps:.2f}/step); "
          f"first yield at {stamps[0]*1e3:.1f} ms, gaps median {np.median(d[1:]):.2f} max {d[1:].max():.2f}, last three {d[-3:].round(2)}")
# host cost of one search_batch call (no waiting): launch only
torch.cuda.synchronize()
t0 = time.perf_counter()
hq.search_batch(db, q, k)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"search_batch host time for one call on an idle GPU: {(t1-t0)*1e3:.2f} ms")
