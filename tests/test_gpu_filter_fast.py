"""GPU: the fast filter (bit planes + per-query cascade) against the exact per-level path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hq():
    import hilbert_quantization_b200 as m
    return m


def _data(rng, N, D, Q, positive):
    if positive:        # every row passes every threshold -> the ratio cut (exact select) decides at each level
        db = (rng.random((N, D)) + 0.25).astype(np.float32)
        qs = (rng.random((Q, D)) + 0.25).astype(np.float32)
    else:
        db = rng.standard_normal((N, D)).astype(np.float32)
        qs = rng.standard_normal((Q, D)).astype(np.float32)
    db[N // 3] = db[5]
    db[N - 2] = db[5]
    qs[0] = db[5]
    return db, qs


@pytest.mark.parametrize("N,D,Q,positive", [(3000, 1536, 20, False), (3000, 1536, 20, True), (5000, 768, 33, False),
                                            (5000, 768, 33, True), (4097, 1024, 9, True), (2000, 256, 130, False),
                                            (70000, 1536, 6, False), (70000, 768, 6, True), (1500, 4096, 7, True)])
@pytest.mark.parametrize("impl", ["fast", "fast_fp32"])        # tensor-core (tf32 hi/lo split) and CUDA-core threshold pass
def test_fast_filter_equals_exact_filter(hq, N, D, Q, positive, impl):
    from hilbert_quantization_b200.search import FilterTrace, unpack_mask
    rng = np.random.default_rng(N + D + Q)
    db, qs = _data(rng, N, D, Q, positive)
    d = hq.EmbeddingDatabase(db)
    assert d.fast_filter_ok
    tf, te = FilterTrace([], [], []), FilterTrace([], [], [])
    if impl == "fast" and D <= 1536:
        assert d.tc_packed is not None           # these layouts fit the 128-float packed operand
    i_f, s_f, m_f = hq.search_batch(d, qs, 10, return_mask=True, filter_impl=impl, trace=tf)
    i_e, s_e, m_e = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="exact", trace=te)
    a_f, a_e = unpack_mask(m_f, N), unpack_mask(m_e, N)
    same = (a_f == a_e).all(axis=1)
    # the two paths evaluate the threshold in different (equally valid) fp32 forms: allow one
    # borderline query, and only single-row differences in it
    assert same.sum() >= Q - 1, f"{(~same).sum()} queries differ"
    for j in np.nonzero(~same)[0]:
        assert (a_f[j] != a_e[j]).sum() <= 2
    ok = torch.from_numpy(same).cuda()
    assert torch.equal(i_f[ok], i_e[ok]) and torch.equal(s_f[ok], s_e[ok])
    for l in range(len(tf.n_out)):
        assert torch.equal(tf.n_out[l][ok], te.n_out[l][ok])
    if positive:
        assert int(tf.n_out[0][0]) == max(1, int(N * 0.3))


@pytest.mark.parametrize("impl", ["fast", "fast_fp32"])
@pytest.mark.parametrize("N,D,positive", [(3000, 768, False), (5000, 1536, False), (3000, 768, True)])
def test_fast_filter_handles_rows_with_short_index_rows(hq, N, D, positive, impl):
    """Rows whose index row ends in an exact zero (stored length < structural length) make the reference take the
    query norm over the shorter prefix: the fast filter scores them as exceptions and must still equal the exact path."""
    from hilbert_quantization_b200.search import FilterTrace, unpack_mask
    rng = np.random.default_rng(N + D)
    db, qs = _data(rng, N, D, 24, positive)
    tail = D - D // 12 if D == 768 else D - 64          # zero the cells of the last live finest-level block
    for r in (17, 18, N // 2, N - 1):
        db[r, tail:] = 0.0
    db[40, D // 2:] = 0.0                               # a row that loses several trailing blocks (and coarser levels)
    qs[3] = db[17] + 0.01 * rng.standard_normal(D).astype(np.float32)
    qs[3, tail:] = db[17, tail:] + 0.05                 # dense query next to an exceptional row
    d = hq.EmbeddingDatabase(db)
    assert d.fast_filter_ok and d.exc_rows.numel() >= 5
    tf, te = FilterTrace([], [], []), FilterTrace([], [], [])
    i_f, s_f, m_f = hq.search_batch(d, qs, 10, return_mask=True, filter_impl=impl, trace=tf)
    i_e, s_e, m_e = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="exact", trace=te)
    a_f, a_e = unpack_mask(m_f, N), unpack_mask(m_e, N)
    same = (a_f == a_e).all(axis=1)
    assert same.sum() >= len(qs) - 1, f"{(~same).sum()} queries differ"
    exc = d.exc_rows.cpu().numpy()
    assert np.array_equal(a_f[:, exc], a_e[:, exc])     # the exceptional rows themselves are decided identically
    ok = torch.from_numpy(same).cuda()
    assert torch.equal(i_f[ok], i_e[ok]) and torch.equal(s_f[ok], s_e[ok])
    for l in range(len(tf.n_out)):
        assert torch.equal(tf.n_out[l][ok], te.n_out[l][ok])


@pytest.mark.parametrize("N,D,Q,offset", [(60000, 1536, 24, 0.03), (60000, 1536, 24, 0.04), (200000, 1536, 8, 0.03)])
def test_list_cascade_when_the_later_ratio_cuts_bind(hq, N, D, Q, offset):
    """N(offset, 1) data: the fine level still decides by its threshold (so the streaming list cascade runs, not the
    per-query fallback) while the coarse levels see similar block means for every row, so MORE rows pass their thresholds
    than the ratio caps allow and the exact selection at the level-1 / level-2 cuts decides.  Fast == exact filter."""
    from hilbert_quantization_b200.search import FilterTrace, unpack_mask
    rng = np.random.default_rng(N + D + Q)
    db = (rng.standard_normal((N, D)) + offset).astype(np.float32)
    qs = (rng.standard_normal((Q, D)) + offset).astype(np.float32)
    db[N // 3] = db[5]
    qs[0] = db[5]
    d = hq.EmbeddingDatabase(db)
    assert d.fast_filter_ok and d.tc_packed is not None
    tf, te = FilterTrace([], [], []), FilterTrace([], [], [])
    i_f, s_f, m_f = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="fast", trace=tf, rerank="bf16")
    i_e, s_e, m_e = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="exact", trace=te, rerank="bf16")
    L = len(te.n_out)
    n_pass = [te.n_pass[l].cpu().numpy() for l in range(L)]
    n_out = [te.n_out[l].cpu().numpy() for l in range(L)]
    # the scenario this test is about: level 0 decided by the threshold for a good share of the queries (those run the
    # list cascade; the others take the per-query fallback), the last level's cap binds for most queries
    assert (n_out[0] == n_pass[0]).mean() > 0.3, (n_out[0], n_pass[0])
    assert (n_out[L - 1] < n_pass[L - 1]).mean() > 0.5, (n_out[L - 1], n_pass[L - 1])
    a_f, a_e = unpack_mask(m_f, N), unpack_mask(m_e, N)
    same = (a_f == a_e).all(axis=1)
    assert same.sum() >= Q - 2, f"{(~same).sum()} queries differ"
    for j in np.nonzero(~same)[0]:
        assert (a_f[j] != a_e[j]).sum() <= 2
    ok = torch.from_numpy(same).cuda()
    assert torch.equal(i_f[ok], i_e[ok]) and torch.equal(s_f[ok], s_e[ok])
    for l in range(L):
        assert torch.equal(tf.n_out[l][ok].to(torch.int64), te.n_out[l][ok].to(torch.int64))


def test_fast_filter_falls_back_on_sparse_data(hq):
    rng = np.random.default_rng(0)
    db = rng.standard_normal((6000, 768)).astype(np.float32)
    db[: 4500, 700:] = 0.0                   # most rows end in a zero run: more exceptions than the fast path takes
    d = hq.EmbeddingDatabase(db)
    assert not d.fast_filter_ok
    with pytest.raises(ValueError, match="fast filter"):
        hq.search_batch(d, db[:3], 5, filter_impl="fast")
    ids, _ = hq.search_batch(d, db[:3], 5)   # auto -> exact path
    assert ids.shape == (3, 5)


@pytest.mark.parametrize("N,D,Q,positive", [(3000, 1536, 12, False), (4097, 768, 9, True)])
def test_global_scope_equals_exact_filter_on_one_shard(hq, N, D, Q, positive):
    """filter_scope="global" with a single shard is the reference's single list: same survivors, ids and
    scores as the exact per-level path (the distributed selection itself is covered over gloo on CPU)."""
    from hilbert_quantization_b200.search import FilterTrace, unpack_mask
    rng = np.random.default_rng(N + Q)
    db, qs = _data(rng, N, D, Q, positive)
    d = hq.EmbeddingDatabase(db)
    tg, te = FilterTrace([], [], []), FilterTrace([], [], [])
    i_g, s_g, m_g = hq.search_batch(d, qs, 10, return_mask=True, filter_scope="global", trace=tg)
    i_e, s_e, m_e = hq.search_batch(d, qs, 10, return_mask=True, filter_impl="exact", trace=te)
    assert np.array_equal(unpack_mask(m_g, N), unpack_mask(m_e, N))
    assert torch.equal(i_g, i_e) and torch.equal(s_g, s_e)
    for l in range(len(te.n_out)):
        assert torch.equal(tg.n_out[l].to(torch.int64), te.n_out[l].to(torch.int64))


def test_sparse_query_is_searched_through_the_exact_path(hq):
    """A query whose index rows are shorter than the structural length cannot use the fast filter: the batch
    is redone through the exact path (the check is read back after the speculative launches)."""
    rng = np.random.default_rng(5)
    db = rng.standard_normal((2000, 768)).astype(np.float32)
    qs = rng.standard_normal((4, 768)).astype(np.float32)
    qs[2, 600:] = 0.0
    d = hq.EmbeddingDatabase(db)
    assert d.fast_filter_ok
    i_a, s_a = hq.search_batch(d, qs, 7)
    i_e, s_e = hq.search_batch(d, qs, 7, filter_impl="exact")
    assert torch.equal(i_a, i_e) and torch.equal(s_a, s_e)
    with pytest.raises(ValueError, match="fast filter"):
        hq.search_batch(d, qs, 7, filter_impl="fast")
