"""GPU: the tcgen05 rerank (bf16 shortlist + exact fp32 re-score) against the exact fp32 path
and the oracle.  Returned scores are fp32-exact, so ids must match and scores agree to 1e-6."""
import numpy as np
import pytest
import torch

from oracle import hilbert_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hq():
    import hilbert_quantization_b200 as m
    return m


@pytest.mark.parametrize("N,D,Q,k", [(1000, 768, 5, 10), (4096, 1536, 128, 10), (5000, 1536, 130, 10), (777, 256, 33, 5),
                                     (20000, 1024, 300, 10), (3000, 1000, 17, 20), (256, 64, 1, 10), (300, 1536, 4, 10)])
def test_tc_rerank_matches_exact_path(hq, N, D, Q, k):
    rng = np.random.default_rng(N + D + Q)
    db = rng.standard_normal((N, D)).astype(np.float32)
    db[N // 2] = db[3]
    db[7] = 0.0                                                 # zero-norm row
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    qs[0] = db[3]
    d = hq.EmbeddingDatabase(db)
    for use_filter in (False, True):
        i32, s32 = hq.search_batch(d, qs, k, use_filter=use_filter, rerank="f32")
        itc, stc = hq.search_batch(d, qs, k, use_filter=use_filter, rerank="bf16")
        assert torch.equal(itc, i32), (use_filter, (itc != i32).sum().item())
        valid = i32 >= 0
        assert (stc[valid] - s32[valid]).abs().max().item() <= 1e-6
        assert torch.equal(stc[~valid], s32[~valid])
    assert itc[0, 0].item() == 3 and itc[0, 1].item() == N // 2     # exact ties resolve to the lower id
    iu, su = hq.search_batch(d, qs[: min(Q, 4)], k, use_filter=False, rerank="bf16")
    for j in range(min(Q, 4)):
        iw, sw = O.topk_stable(np.arange(N), O.cosine01(qs[j], db), k)
        assert list(iu[j].cpu().numpy()) == list(iw)
        assert np.abs(su[j].cpu().numpy() - sw).max() < 5e-7


@pytest.mark.parametrize("N,D,Q", [(5000, 1536, 1), (20000, 768, 4), (3001, 250, 3), (70000, 1024, 2)])
def test_sparse_rerank_equals_tensor_core_path(hq, N, D, Q):
    """Latency path (Q <= 4, what rerank="auto" picks): only the filter's survivors are scored, in exact fp32.
    Ids and scores must be IDENTICAL to the tensor-core path (which re-scores its shortlist with the same arithmetic)."""
    rng = np.random.default_rng(N + Q)
    db = rng.standard_normal((N, D)).astype(np.float32)
    db[N // 2] = db[3]
    db[7] = 0.0
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    qs[0] = db[3] + 0.001 * rng.standard_normal(D).astype(np.float32)
    d = hq.EmbeddingDatabase(db)
    i_auto, s_auto = hq.search_batch(d, qs, 10)                       # auto -> sparse
    i_sp, s_sp = hq.search_batch(d, qs, 10, rerank="sparse")
    i_tc, s_tc = hq.search_batch(d, qs, 10, rerank="bf16")
    assert torch.equal(i_auto, i_sp) and torch.equal(s_auto, s_sp)
    i_sb, s_sb = hq.search_batch(d, qs, 10, rerank="sparse_bf16")     # bf16 shortlist + exact re-score, proven or redone
    assert torch.equal(i_sb, i_sp) and torch.equal(s_sb, s_sp)
    assert torch.equal(i_sp, i_tc), (i_sp, i_tc)
    assert torch.equal(s_sp, s_tc)
    assert set(i_sp[0, :2].tolist()) == {3, N // 2}
    for j in range(Q):
        iw, sw = O.progressive_search(qs[j], db, d.n, 10)
        assert list(i_sp[j, : len(iw)].cpu().numpy()) == list(iw)


def test_tc_rerank_sharded_id_base(hq):
    rng = np.random.default_rng(3)
    db = rng.standard_normal((2048, 512)).astype(np.float32)
    qs = rng.standard_normal((9, 512)).astype(np.float32)
    full = hq.EmbeddingDatabase(db)
    a = hq.EmbeddingDatabase(db[:1000], id_base=0)
    b = hq.EmbeddingDatabase(db[1000:], id_base=1000)
    ia, sa = hq.search_batch(a, qs, 10, use_filter=False)
    ib, sb = hq.search_batch(b, qs, 10, use_filter=False)
    from hilbert_quantization_b200.distributed import merge_topk_host
    mi, ms = merge_topk_host(np.stack([ia.cpu().numpy(), ib.cpu().numpy()]), np.stack([sa.cpu().numpy(), sb.cpu().numpy()]), 10)
    fi, fs = hq.search_batch(full, qs, 10, use_filter=False)
    assert np.array_equal(mi, fi.cpu().numpy())


def test_tc_rerank_lists_zero_norm_rows_last(hq):
    """The tensor-core operand holds unit rows, zero-norm rows are appended from their id list: they score exactly 0.0
    (rag/search/engine.py:640-643) and follow every other survivor in id order."""
    rng = np.random.default_rng(11)
    db = rng.standard_normal((40, 768)).astype(np.float32)
    zero = [2, 5, 17, 30, 31, 39]
    db[zero] = 0.0
    qs = rng.standard_normal((3, 768)).astype(np.float32)
    for n_rows, k in ((8, 10), (40, 10), (40, 20), (20, 16)):
        d = hq.EmbeddingDatabase(db[:n_rows])
        i32, s32 = hq.search_batch(d, qs, k, use_filter=False, rerank="f32")
        itc, stc = hq.search_batch(d, qs, k, use_filter=False, rerank="bf16")
        assert torch.equal(itc, i32), (n_rows, k)
        assert torch.equal(stc == 0, s32 == 0)
        for j in range(3):
            iw, sw = O.topk_stable(np.arange(n_rows), O.cosine01(qs[j], db[:n_rows]), k)
            got = itc[j].cpu().numpy()
            assert list(got[: len(iw)]) == list(iw)


@pytest.mark.parametrize("N,D", [(1000, 1536), (333, 250), (64, 1537), (5, 8), (2049, 771)])
def test_unit_row_bf16_operand_is_exact(hq, N, D):
    """the rerank GEMM's database operand: c / |c| in fp32 (IEEE division), rounded to nearest-even bf16, zero rows and the
    padding columns zero -- bit for bit against torch, for row lengths that do and do not fill the 8-element chunks"""
    from hilbert_quantization_b200.search import row_norms, to_bf16
    g = torch.Generator(device="cuda").manual_seed(N * 7 + D)
    x = torch.randn((N, D), device="cuda", generator=g)
    x[N // 2] = 0.0
    norms = row_norms(x)
    got = to_bf16(x, norms)
    pitch = (D + 7) // 8 * 8
    assert got.shape == (N, pitch) and got.dtype == torch.bfloat16
    want = torch.zeros((N, pitch), dtype=torch.bfloat16, device="cuda")
    want[:, :D] = torch.where(norms[:, None] > 0, x / norms[:, None], torch.zeros_like(x)).to(torch.bfloat16)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    sub = x[:, : D - D % 4 - 4] if D > 8 else x                      # a row-pitched view: stride > width
    if sub.shape[1] > 0 and not sub.is_contiguous():
        n2 = row_norms(sub)
        got2 = to_bf16(sub, n2)
        w2 = torch.where(n2[:, None] > 0, sub / n2[:, None], torch.zeros_like(sub)).to(torch.bfloat16)
        assert torch.equal(got2[:, : sub.shape[1]].view(torch.int16), w2.view(torch.int16)) and not got2[:, sub.shape[1]:].any()


@pytest.mark.parametrize("N,D,Q", [(50000, 1536, 1), (20000, 768, 3), (30000, 1024, 64), (4097, 250, 2)])
def test_search_graph_replays_the_same_search(hq, N, D, Q):
    """SearchGraph = search_batch captured once as a CUDA graph: identical ids and scores on every replay, for the sparse
    (few queries) and the tensor-core rerank, with the fast and (odd D) the exact filter; a batch with a sparse query index row
    is answered through the exact path like search_batch does."""
    g = torch.Generator(device="cuda").manual_seed(N + Q)
    db = torch.randn((N, D), device="cuda", generator=g)
    d = hq.EmbeddingDatabase(db)
    sg = hq.SearchGraph(d, Q, 10)
    for rep in range(3):
        qs = torch.randn((Q, D), device="cuda", generator=g)
        qs[0] = db[rep * 7 + 1]
        ids, sc = sg.search(qs)
        ids, sc = ids.clone(), sc.clone()
        want_i, want_s = hq.search_batch(d, qs, 10)
        assert torch.equal(ids, want_i) and torch.equal(sc, want_s)
        assert ids[0, 0].item() == rep * 7 + 1
    ids, sc = sg.search(qs.cpu().numpy())                              # host queries
    assert torch.equal(ids, want_i) and torch.equal(sc, want_s)
    if D % 64 == 0:
        sparse_q = qs.clone()
        sparse_q[0, D - 64:] = 0.0                                     # last level-0 block mean exactly 0: a shortened index row
        ids, sc = sg.search(sparse_q)
        want_i, want_s = hq.search_batch(d, sparse_q, 10)
        assert torch.equal(ids, want_i) and torch.equal(sc, want_s)
    with pytest.raises(ValueError):
        sg.search(torch.randn((Q + 1, D), device="cuda"))
    # a larger eager batch makes the shard replace its filter scratch; the graph keeps the allocation it captured
    big = torch.randn((4 * Q + 200, D), device="cuda", generator=g)
    hq.search_batch(d, big, 10)
    junk = [torch.full((1 << 22,), 7.0, device="cuda") for _ in range(8)]          # would land in a freed scratch
    qs = torch.randn((Q, D), device="cuda", generator=g)
    ids, sc = sg.search(qs)
    want_i, want_s = hq.search_batch(d, qs, 10)
    assert torch.equal(ids, want_i) and torch.equal(sc, want_s) and all(bool((j == 7.0).all()) for j in junk)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_search_stream_pipelines_host_batches(hq, depth):
    """search_stream = search_batch per host batch with the copies overlapped: same ids / scores, in order, for more batches
    than slots; results are pinned host tensors"""
    g = torch.Generator(device="cuda").manual_seed(depth)
    db = torch.randn((40000, 1024), device="cuda", generator=g)
    d = hq.EmbeddingDatabase(db)
    batches = [torch.randn((96, 1024), generator=torch.Generator().manual_seed(100 + i)).pin_memory() for i in range(7)]
    for b in batches:
        b[0] = db[11].cpu()
    got = [(i.clone(), s.clone()) for i, s in hq.search_stream(d, batches, 10, depth=depth)]
    assert len(got) == len(batches)
    for b, (ids, sc) in zip(batches, got):
        want_i, want_s = hq.search_batch(d, b.cuda(), 10)
        assert not ids.is_cuda and torch.equal(ids, want_i.cpu()) and torch.equal(sc, want_s.cpu())
        assert ids[0, 0].item() == 11
    # a post hook (what the sharded search uses for its all-gather merge) and numpy batches
    got2 = list(hq.search_stream(d, [b.numpy() for b in batches[:2]], 10, depth=depth, post=lambda i, s: (i + 5, s)))
    assert torch.equal(got2[-1][0], got[1][0] + 5)
    with pytest.raises(ValueError):
        list(hq.search_stream(d, [batches[0], batches[1][:5]], 10, depth=depth))


# ------------------------------------------------------------------------------------------------
# shortlist-sufficiency guard (hq_rerank_topk_unit_bf16): the bf16 accumulator only PROPOSES rows; the
# returned top-k is proven per query, and queries that cannot be proven are re-scored exactly
# ------------------------------------------------------------------------------------------------
def _check_topk_modulo_near_ties(ids, sc, qs, db, alive, k, tol=5e-7):
    """ids / scores are a correct top-k of the eligible rows: the score list equals the fp64 oracle's sorted top-k scores within
    `tol`, every returned row is eligible and its own fp64 score is the returned one (rows whose scores differ by less than the
    fp32 resolution may legitimately swap places)."""
    for j in range(len(qs)):
        rows = np.nonzero(alive[j])[0]
        s64 = O.cosine01(qs[j], db[rows])
        want = np.sort(s64)[::-1][:k]
        m = len(want)
        assert (ids[j, :m] >= 0).all() and (ids[j, m:] == -1).all(), j
        if m == 0:
            continue
        assert np.abs(sc[j, :m] - want).max() <= tol, (j, np.abs(sc[j, :m] - want).max())
        pos = {int(r): i for i, r in enumerate(rows)}
        own = np.array([s64[pos[int(i)]] for i in ids[j, :m]])           # KeyError = a row that is not eligible
        assert np.abs(own - sc[j, :m]).max() <= tol, j
        assert len(set(ids[j, :m].tolist())) == m


def _stats(guard_list):
    g = torch.stack([t[:2] for t in guard_list]).sum(0).tolist()
    return int(g[0]), int(g[1])


def test_guard_passes_on_random_data_and_counts_rescored_rows(hq):
    rng = np.random.default_rng(11)
    N, D, Q, k = 60000, 1536, 64, 10
    db = rng.standard_normal((N, D)).astype(np.float32)
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    d = hq.EmbeddingDatabase(db)
    assert 0.0 < d.dc_max < 2.0 ** -8                       # |bf16(c/|c|) - c/|c|| <= 2^-8 |c/|c||
    stats = []
    itc, stc = hq.search_batch(d, qs, k, use_filter=False, rerank="bf16", guard_stats=stats)
    i32, s32 = hq.search_batch(d, qs, k, use_filter=False, rerank="f32")
    assert torch.equal(itc, i32) and (stc - s32).abs().max().item() <= 1e-6
    flagged, rescored = _stats(stats)
    assert flagged == 0, flagged                            # no query needed the exact fallback
    assert Q * 16 <= rescored <= Q * 64, rescored           # 16 .. 64 rows per query were re-scored exactly


@pytest.mark.parametrize("use_filter", [False, True])
def test_guard_catches_near_duplicate_clusters(hq, use_filter):
    """ADVERSARIAL: 50 000 rows within ~1e-4 cosine of each other.  Hundreds of rows lie inside the bf16 error band of the
    10th score, so a fixed shortlist of 16 drops true top-10 rows; the guard must flag those queries and the exact fallback
    must return the exact path's ids and scores (rag/search/engine.py:622-660, :512)."""
    rng = np.random.default_rng(5)
    N, D, Q, k = 50000, 768, 24, 10
    base = rng.standard_normal(D).astype(np.float32)
    base /= np.linalg.norm(base)
    db = base[None, :] + (1.0e-2 / np.sqrt(D)) * rng.standard_normal((N, D)).astype(np.float32)   # cos(row, row') ~ 1 - 1e-4
    db[100:110] = 0.5 * base[None, :] + 0.5 * rng.standard_normal((10, D)).astype(np.float32) / np.sqrt(D)   # a few ordinary rows
    qs = base[None, :] + (1.0e-2 / np.sqrt(D)) * rng.standard_normal((Q, D)).astype(np.float32)
    qs[Q // 2:] = rng.standard_normal((Q - Q // 2, D)).astype(np.float32)       # half the queries see the cluster from far away
    d = hq.EmbeddingDatabase(db)
    stats = []
    itc, stc, mask = hq.search_batch(d, qs, k, use_filter=use_filter, rerank="bf16", guard_stats=stats, return_mask=True)
    from hilbert_quantization_b200.search import unpack_mask
    alive = unpack_mask(mask, N) if use_filter else np.ones((Q, N), dtype=bool)
    _check_topk_modulo_near_ties(itc.cpu().numpy(), stc.cpu().numpy(), qs, db, alive, k)
    if use_filter:          # the sparse fp32 rerank scores every survivor with the arithmetic of the re-score: identical results
        isp, ssp = hq.search_batch(d, qs, k, rerank="sparse")
        assert torch.equal(itc, isp), (itc != isp).sum().item()
        assert torch.equal(stc, ssp)
    flagged, _ = _stats(stats)
    assert flagged >= 1, "the adversarial cluster should defeat the 64-row shortlist for at least one query"
    # and the unguarded shortlist really would have been wrong: the exact top-10 is not inside the bf16 top-16
    q16 = torch.from_numpy(qs[:1]).cuda().bfloat16().float()
    acc = (q16 @ d.emb_bf16[:, :D].float().T)[0]
    top16 = set(acc.topk(16).indices.tolist())
    exact_top = set(np.argsort(-O.cosine01(qs[0], db), kind="stable")[:k].tolist())
    assert not exact_top <= top16


def test_guard_zero_query_and_all_ties(hq):
    rng = np.random.default_rng(2)
    N, D = 3000, 256
    db = rng.standard_normal((N, D)).astype(np.float32)
    db[50:2050] = db[50]                                    # 2000 identical rows: every one of them ties
    qs = np.stack([np.zeros(D, np.float32), db[50], rng.standard_normal(D).astype(np.float32)])
    d = hq.EmbeddingDatabase(db)
    itc, stc = hq.search_batch(d, qs, 10, use_filter=False, rerank="bf16")
    i32, s32 = hq.search_batch(d, qs, 10, use_filter=False, rerank="f32")
    assert torch.equal(itc, i32) and (stc - s32).abs().max().item() <= 1e-6
    assert itc[0].tolist() == list(range(10))               # zero query: every score is 0.0, lowest ids win
    assert itc[1].tolist() == list(range(50, 60))


@pytest.mark.parametrize("N,D,Q,k", [(30000, 768, 40, 10), (9000, 1536, 130, 20), (2000, 250, 3, 5)])
def test_bf16_only_database_scores_the_stored_rows_exactly(hq, N, D, Q, k):
    """EmbeddingDatabase.from_chunks keeps no fp32 rows: the stored rows are the bf16 unit rows, the returned scores are the
    exact (cos + 1) / 2 of the query with the STORED row (<= 5e-7), ids equal the oracle's top-k over the stored rows, and
    the scores stay within the stated bf16 tolerance (2e-3) of the fp32 database's."""
    rng = np.random.default_rng(N + D)
    db = rng.standard_normal((N, D)).astype(np.float32)
    db[5] = 0.0
    db[N // 3] = db[9]
    qs = rng.standard_normal((Q, D)).astype(np.float32)
    qs[0] = db[9]
    chunks = [torch.from_numpy(db[s:s + 7000]).cuda() for s in range(0, N, 7000)]
    lean = hq.EmbeddingDatabase.from_chunks(iter(chunks), N, D)
    full = hq.EmbeddingDatabase(db)
    assert lean.emb is None and torch.equal(lean.idx, full.idx)
    if D % 8 == 0:
        assert torch.equal(lean.emb_bf16, full.emb_bf16)
    stored = lean.emb_bf16[:, :D].float().cpu().numpy()
    for use_filter in (False, True):
        ids, sc = hq.search_batch(lean, qs, k, use_filter=use_filter)
        idf, scf = hq.search_batch(full, qs, k, use_filter=use_filter, return_mask=False)
        ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
        if use_filter:
            _, _, mask = hq.search_batch(lean, qs, k, return_mask=True)
            from hilbert_quantization_b200.search import unpack_mask
            alive = unpack_mask(mask, N)
        for j in range(min(Q, 12)):
            rows = np.nonzero(alive[j])[0] if use_filter else np.arange(N)
            iw, sw = O.topk_stable(rows, O.cosine01(qs[j], stored[rows]), k)
            m = len(iw)
            assert list(ids[j, :m]) == list(iw), (use_filter, j)
            assert np.abs(sc[j, :m] - sw).max() < 5e-7
        valid = (ids >= 0) & (idf.cpu().numpy() == ids)
        assert np.abs(sc[valid] - scf.cpu().numpy()[valid]).max() < 2e-3
    with pytest.raises(ValueError, match="bf16-only"):
        hq.search_batch(lean, qs, k, rerank="f32")


def _unit_rows(lib, dev, check, db, norms):
    """bf16 unit rows + dc_max of a device matrix (what EmbeddingDatabase keeps for the tensor-core rerank)."""
    N, D = db.shape
    pitch = (D + 7) // 8 * 8
    unit = torch.zeros((N, pitch), dtype=torch.bfloat16, device="cuda")
    check(lib.hq_to_bf16_unit(dev.ptr(db), N, D, D, dev.ptr(norms), dev.ptr(unit), pitch, dev.stream_ptr()))
    worst = torch.zeros(1, dtype=torch.float32, device="cuda")
    check(lib.hq_bf16_unit_error_max(dev.ptr(db), N, D, D, dev.ptr(norms), dev.ptr(unit), pitch, dev.ptr(worst), dev.stream_ptr()))
    return unit, pitch, float(worst.item())


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("N,D,Q,k,density", [(50_000, 1536, 1, 10, 0.06), (9_999, 768, 4, 32, 0.3), (4_097, 2048, 2, 1, 0.5),
                                             (3_000, 4096, 3, 10, 0.02), (70_001, 100, 1, 10, 0.9), (1_000, 1024, 2, 20, 0.004),
                                             (6_000, 1536, 1, 10, 0.97), (2_500, 8192, 2, 5, 0.6)])
def test_fused_sparse_rerank_topk_equals_the_two_step_path(hq, N, D, Q, k, density, bf16):
    """hq_rerank_sparse_topk (rows staged by bulk copies, per-warp lists, last-CTA merge; with `bf16` the survivors are ranked
    by their bf16 unit rows first and the best 32 re-scored exactly, unproven queries redone over the fp32 rows) against
    hq_rerank_scores_sparse_f32 + hq_topk_from_scores_chunked on random survivor masks: ids and scores identical, with
    duplicate rows (exact ties -> lower id), zero rows, fewer than k survivors and a non-zero id_base."""
    from hilbert_quantization_b200 import _device as dev
    from hilbert_quantization_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(N + D + k)
    db = torch.randn((N, D), device="cuda", generator=g)
    db[N // 2] = db[3]
    db[N - 1] = db[3]
    db[7] = 0.0
    qs = torch.randn((Q, D), device="cuda", generator=g)
    qs[0] = db[3]
    norms, qn = db.norm(dim=1), qs.norm(dim=1)
    check(lib.hq_row_norms(dev.ptr(db), N, D, D, dev.ptr(norms), dev.stream_ptr()))
    check(lib.hq_row_norms(dev.ptr(qs), Q, D, D, dev.ptr(qn), dev.stream_ptr()))
    words = (N + 31) // 32
    alive = torch.rand((Q, words * 32), device="cuda", generator=g) < density
    alive[:, [3, 7, N // 2, N - 1]] = True
    alive[:, N:] = True                                            # bits beyond N must be ignored
    if Q > 1:
        alive[Q - 1] = False
        alive[Q - 1, :5] = True                                    # fewer than k survivors (k > 5)
        alive[Q - 1, N:] = True
    shifts = torch.arange(32, device="cuda", dtype=torch.int64)
    mask = ((alive.view(Q, words, 32).to(torch.int64) << shifts).sum(dim=2) & 0xffffffff).to(torch.int64)
    mask = torch.where(mask >= 2 ** 31, mask - 2 ** 32, mask).to(torch.int32).contiguous()
    id_base = 1_000_000
    assert lib.hq_rerank_sparse_topk_supported(D, D, D, k)
    unit, pitch, dc_max = _unit_rows(lib, dev, check, db, norms) if bf16 else (None, 0, 0.0)
    ids = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    sb = int(lib.hq_rerank_sparse_topk_scratch_bytes(Q, k))
    scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
    for _ in range(2):                                             # the second call finds the counters as the first left them
        check(lib.hq_rerank_sparse_topk(dev.ptr(db), dev.ptr(norms), N, D, D, dev.ptr(unit), pitch, dc_max, dev.ptr(qs), dev.ptr(qn), Q, D,
                                        dev.ptr(mask), words, k, id_base, dev.ptr(ids), dev.ptr(sc), dev.ptr(scratch), sb,
                                        dev.stream_ptr()))
    scores = torch.empty((Q, N), dtype=torch.float32, device="cuda")
    check(lib.hq_rerank_scores_sparse_f32(dev.ptr(db), dev.ptr(norms), N, D, D, dev.ptr(qs), dev.ptr(qn), Q, D, dev.ptr(mask), words,
                                          dev.ptr(scores), N, dev.stream_ptr()))
    ids2 = torch.empty_like(ids)
    sc2 = torch.empty_like(sc)
    tb = max(int(lib.hq_topk_chunked_scratch_bytes(N, Q, k)), 8)
    ts = torch.empty(tb, dtype=torch.uint8, device="cuda")
    check(lib.hq_topk_from_scores_chunked(dev.ptr(scores), N, N, Q, k, id_base, dev.ptr(ids2), dev.ptr(sc2), dev.ptr(ts), tb, dev.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(ids, ids2), (ids, ids2)
    assert torch.equal(sc, sc2)
    if k >= 3:
        assert ids[0, :3].tolist() == [id_base + 3, id_base + N // 2, id_base + N - 1]
    if Q > 1 and k > 5:
        assert (ids[Q - 1, 5:] == -1).all() and (sc[Q - 1, 5:] == -1.0).all()


def test_fused_sparse_rerank_bf16_shortlist_on_a_near_duplicate_cluster(hq):
    """20 000 rows within 1e-4 (cosine) of each other around the query: the bf16 accumulators cannot order them, the shortlist
    check must fail and the fp32 pass behind it must return the exact top-k."""
    from hilbert_quantization_b200 import _device as dev
    from hilbert_quantization_b200._lib import lib, check
    N, D, k = 60_000, 1536, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn(D, device="cuda", generator=g)
    db = torch.randn((N, D), device="cuda", generator=g)
    db[:20_000] = base + 0.01 * torch.randn((20_000, D), device="cuda", generator=g)
    qs = (base + 0.01 * torch.randn(D, device="cuda", generator=g))[None].contiguous()
    norms, qn = torch.empty(N, device="cuda"), torch.empty(1, device="cuda")
    check(lib.hq_row_norms(dev.ptr(db), N, D, D, dev.ptr(norms), dev.stream_ptr()))
    check(lib.hq_row_norms(dev.ptr(qs), 1, D, D, dev.ptr(qn), dev.stream_ptr()))
    unit, pitch, dc_max = _unit_rows(lib, dev, check, db, norms)
    words = (N + 31) // 32
    mask = torch.full((1, words), -1, dtype=torch.int32, device="cuda")
    out = []
    for u in (unit, None):
        ids = torch.empty((1, k), dtype=torch.int64, device="cuda")
        sc = torch.empty((1, k), dtype=torch.float32, device="cuda")
        sb = int(lib.hq_rerank_sparse_topk_scratch_bytes(1, k))
        scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
        check(lib.hq_rerank_sparse_topk(dev.ptr(db), dev.ptr(norms), N, D, D, dev.ptr(u), pitch if u is not None else 0, dc_max,
                                        dev.ptr(qs), dev.ptr(qn), 1, D, dev.ptr(mask), words, k, 0, dev.ptr(ids), dev.ptr(sc),
                                        dev.ptr(scratch), sb, dev.stream_ptr()))
        out.append((ids, sc))
    torch.cuda.synchronize()
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    exact = ((db[:20_000].double() @ qs[0].double()) / (db[:20_000].double().norm(dim=1) * qs[0].double().norm()) + 1) / 2
    want = torch.topk(exact, k).indices
    assert set(out[0][0][0].tolist()) == set(want.tolist())


def test_fused_sparse_rerank_topk_with_all_scores_tied(hq):
    """Every surviving row identical: all candidates tie, the merge must fall back to ranking by id."""
    from hilbert_quantization_b200 import _device as dev
    from hilbert_quantization_b200._lib import lib, check
    N, D, k = 40_000, 768, 10
    row = torch.randn(D, device="cuda")
    db = row.repeat(N, 1).contiguous()
    qs = row[None].clone()
    norms, qn = torch.empty(N, device="cuda"), torch.empty(1, device="cuda")
    check(lib.hq_row_norms(dev.ptr(db), N, D, D, dev.ptr(norms), dev.stream_ptr()))
    check(lib.hq_row_norms(dev.ptr(qs), 1, D, D, dev.ptr(qn), dev.stream_ptr()))
    words = (N + 31) // 32
    mask = torch.full((1, words), -1, dtype=torch.int32, device="cuda")
    mask[0, 0] = -4                                                 # rows 0 and 1 are dead
    ids = torch.empty((1, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((1, k), dtype=torch.float32, device="cuda")
    sb = int(lib.hq_rerank_sparse_topk_scratch_bytes(1, k))
    scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
    check(lib.hq_rerank_sparse_topk(dev.ptr(db), dev.ptr(norms), N, D, D, None, 0, 0.0, dev.ptr(qs), dev.ptr(qn), 1, D, dev.ptr(mask), words,
                                    k, 0, dev.ptr(ids), dev.ptr(sc), dev.ptr(scratch), sb, dev.stream_ptr()))
    torch.cuda.synchronize()
    assert ids[0].tolist() == list(range(2, 2 + k))
    assert (sc[0] == sc[0, 0]).all()
