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
