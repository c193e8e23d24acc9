"""Randomised cross-check of the search paths on one GPU (not part of the test suite: ~1 min):
fast filter (tcgen05 + list cascade) vs exact filter, tensor-core / sparse rerank vs the exact fp32 rerank, on random
shapes incl. odd N, D, Q, duplicates, zero rows, positive data (ratio cuts bind) and small shards.

    python tools/stress.py [seed] [cases]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import hilbert_quantization_b200 as hq
from hilbert_quantization_b200.search import unpack_mask

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(seed)
bad = 0
for c in range(cases):
    D = int(rng.choice([64, 250, 256, 768, 1000, 1024, 1536, 1537, 3000, 4096]))
    N = int(rng.choice([1, 2, 31, 63, 64, 65, 200, 777, 4096, 5001, 33333, 150000, 420000]))
    if N * D > 6e8:
        N = int(6e8 // D)
    Q = int(rng.choice([1, 2, 3, 4, 5, 17, 127, 128, 129, 300]))
    k = int(rng.choice([1, 5, 10, 20]))
    positive = bool(rng.random() < 0.4)
    g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
    db = torch.randn((N, D), device="cuda", generator=g)
    qs = torch.randn((Q, D), device="cuda", generator=g)
    if positive:
        db, qs = db.abs() + 0.25, qs.abs() + 0.25
    if N > 8:
        db[N // 2] = db[3]
        db[5] = 0.0
        qs[0] = db[3]
    d = hq.EmbeddingDatabase(db)
    tag = f"case {c}: N={N} D={D} Q={Q} k={k} positive={positive} fast_ok={d.fast_filter_ok}"
    try:
        i_e, s_e, m_e = hq.search_batch(d, qs, k, return_mask=True, filter_impl="exact", rerank="f32")
        i_a, s_a, m_a = hq.search_batch(d, qs, k, return_mask=True)                     # auto: fast filter, tc or sparse rerank
        a_e, a_a = unpack_mask(m_e, N), unpack_mask(m_a, N)
        same = (a_e == a_a).all(axis=1)
        n_diff_rows = int((a_e != a_a).sum())
        ok = torch.from_numpy(same).cuda()
        # Survivor sets: the two filters evaluate the same fp32 threshold test in different (equally valid) forms, so a row
        # whose level score lies within ~2e-7 of a threshold / cut score may fall on either side: with a score density of
        # ~2 per unit that is ~1e-6 * N rows per query and level.  Anything beyond a generous multiple of that is a bug.
        allowed_rows = 2 + int(4e-6 * N * Q * d.num_levels)
        # Top-k of the queries whose survivor sets agree: same id SET; positions may differ only between near-tied scores
        # (two fp32 summation orders of a D-term dot product differ by up to a few 1e-6 on all-positive data).
        tol = 5e-6
        ie, ia, se, sa = i_e[ok].cpu().numpy(), i_a[ok].cpu().numpy(), s_e[ok].cpu().numpy(), s_a[ok].cpu().numpy()
        ids_ok, sc_ok, worst = True, True, 0.0
        for r in range(ie.shape[0]):
            if not np.array_equal(ie[r], ia[r]):
                kth = se[r][ie[r] >= 0].min() if (ie[r] >= 0).any() else 0.0
                for col in np.nonzero(ie[r] != ia[r])[0]:
                    # a swapped / replaced entry must be a near tie with its counterpart (or with the k-th score)
                    if abs(se[r][col] - sa[r][col]) > tol and abs(se[r][col] - kth) > tol:
                        ids_ok = False
            v = ie[r] >= 0
            if v.any():
                diff = np.abs(np.sort(se[r][v]) - np.sort(sa[r][v])).max()
                worst = max(worst, float(diff))
                sc_ok = sc_ok and diff <= tol
        if n_diff_rows > allowed_rows or not ids_ok or not sc_ok:
            bad += 1
            print("MISMATCH", tag, "queries differing:", int((~same).sum()), "rows:", n_diff_rows, "allowed", allowed_rows,
                  "ids_ok", ids_ok, "scores_ok", sc_ok, "worst score diff %.2e" % worst)
        else:
            print("ok", tag, "borderline queries:", int((~same).sum()), "rows:", n_diff_rows, "worst score diff %.1e" % worst)
    except Exception as e:                                                             # noqa: BLE001
        bad += 1
        print("ERROR", tag, repr(e)[:300])
    del d, db, qs
    torch.cuda.empty_cache()
print("stress done:", cases, "cases,", bad, "bad")
sys.exit(1 if bad else 0)
