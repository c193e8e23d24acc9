"""GPU parity at BASELINE.json's FULL sizes (C2, C3, C4 and one C5 shard) through size-independent properties:
round trips, the oracle's permutation applied to whole grids, torch fp32/fp64 recomputation of what the search returns,
the ratio caps of rag/search/engine.py:272-279, and fast filter == exact filter on a sample of the queries.  The oracle
cannot run these sizes in seconds (SURVEY 8c "scale limits"), so the small-size tests carry the value-by-value parity and
these carry indexing (> 2^32 elements in one launch), grid sizing and scratch sizing.  Runs last (file name) so a
failure here cannot hide the parity tests under `pytest -x`."""
import gc

import numpy as np
import pytest
import torch

from oracle import hilbert_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hq():
    import hilbert_quantization_b200 as m
    return m


@pytest.fixture(autouse=True)
def _release_device_memory():
    gc.collect()
    torch.cuda.empty_cache()
    yield
    gc.collect()
    torch.cuda.empty_cache()


def _need_gb(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory, {free >> 30} GB free")


def _unit_rows(rows, dim, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty((rows, dim), dtype=torch.float32, device="cuda")
    step = 1 << 18
    for s in range(0, rows, step):
        e = min(rows, s + step)
        x[s:e] = torch.randn((e - s, dim), generator=g, device="cuda")
        x[s:e] /= x[s:e].norm(dim=1, keepdim=True)
    return x


def _queries(db, Q, seed):
    """half perturbed database rows (row + 0.1 randn, renormalised), half fresh random directions"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn((Q, db.shape[1]), generator=g, device="cuda")
    half = Q // 2
    q[:half] = db[:half] + 0.1 * q[:half]
    return q / q.norm(dim=1, keepdim=True)


# ---------------------------------------------------------------------------------------------------------------
# C3: 10 M x 1024 -> 32 x 32 -> back, ONE launch each way over 1.024e10 elements (> 2^32)
# ---------------------------------------------------------------------------------------------------------------
def test_c3_round_trip_10m_rows_one_launch(hq):
    _need_gb(130)
    N, D, n = 10_000_000, 1024, 32
    x = torch.empty((N, D), dtype=torch.int32, device="cuda")
    step = 1 << 20
    for s in range(0, N, step):                      # every element unique within its row and tagged with its row
        e = min(N, s + step)
        rows = torch.arange(s, e, device="cuda", dtype=torch.int64)[:, None]
        cols = torch.arange(D, device="cuda", dtype=torch.int64)[None, :]
        x[s:e] = ((rows * 2654435761 + cols * 40503) & 0x7FFFFFFF).to(torch.int32)
    x = x.view(torch.float32)                        # 4-byte copies: the bit patterns travel untouched
    m = hq.HilbertCurveMapper()
    grids = m.map_to_2d_batch(x, n)
    assert grids.shape == (N, n, n)
    # the oracle's permutation (core/hilbert_mapper.py:149-161) on the first, the last and a middle block of rows
    xs, ys = O.hilbert_coordinates(n)
    cell = torch.from_numpy((ys.astype(np.int64) * n + xs.astype(np.int64))).cuda()
    for s in (0, N // 2 - 500, N - 1000):
        want = torch.empty((1000, n * n), dtype=torch.int32, device="cuda")
        want[:, cell] = x[s:s + 1000].view(torch.int32)
        assert torch.equal(grids[s:s + 1000].view(torch.int32).reshape(1000, -1), want)
    back = m.map_from_2d_batch(grids)
    del grids
    xi, bi = x.view(torch.int32), back.view(torch.int32)
    for s in range(0, N, 1 << 21):
        assert torch.equal(bi[s:s + (1 << 21)], xi[s:s + (1 << 21)])


# ---------------------------------------------------------------------------------------------------------------
# C4: 494,032,768-value parameter stream -> 30 grids of 4096 x 4096 (last at fill 0.447) + variant-C indices
# ---------------------------------------------------------------------------------------------------------------
def test_c4_parameter_stream_4096_grids(hq):
    _need_gb(12)
    total, n = 494_032_768, 4096
    cells = n * n
    G = -(-total // cells)
    g = torch.Generator(device="cuda").manual_seed(0)
    stream = torch.empty(total, dtype=torch.float32, device="cuda").normal_(0.0, 0.02, generator=g)
    grids, idx = hq.map_parameter_stream(stream, n, variant="C")
    assert grids.shape == (G, n, n) and idx.shape[0] == G
    # whole grids against the oracle's coordinates: grid[y_d, x_d] = stream[d], padding cells stay 0
    xs, ys = O.hilbert_coordinates(n)
    cell = torch.from_numpy(ys.astype(np.int64) * n + xs.astype(np.int64)).cuda()
    for gi in (0, G - 2, G - 1):
        vals = torch.zeros(cells, dtype=torch.float32, device="cuda")
        part = stream[gi * cells:(gi + 1) * cells]
        vals[:part.numel()] = part
        want = torch.empty(cells, dtype=torch.float32, device="cuda")
        want[cell] = vals
        assert torch.equal(grids[gi].reshape(-1).view(torch.int32), want.view(torch.int32))
    back = hq.HilbertCurveMapper().map_from_2d_batch(grids).view(-1)
    assert torch.equal(back[:total].view(torch.int32), stream.view(torch.int32)) and not back[total:].any()
    # finest index row of variant C (64 x 64 sections = runs of 4096 curve positions, rag/.../hierarchical_index_generator.py:
    # 138-178) against fp64 block means of the grid itself, in the section order the oracle states; 3e-7 is the fp32 bound
    # of the small-size tests, the values here are ~N(0, 0.02) / sqrt(4096)
    padded = torch.zeros(G * cells, dtype=torch.float64, device="cuda")
    padded[:total] = stream.double()
    lvl0 = padded.view(G, cells // 4096, 4096).mean(dim=2)
    assert (idx[:, :4096].double() - lvl0).abs().max().item() < 3e-7
    # the whole variant-C index row of the partially filled last grid against the oracle (9 s of NumPy for one 4096^2 grid)
    want = O.index_c_batch_compact(grids[G - 1:G].cpu().numpy())
    assert want.shape[1] == idx.shape[1]
    assert np.abs(idx[G - 1].cpu().numpy().astype(np.float64) - want[0].astype(np.float64)).max() < 3e-7


# ---------------------------------------------------------------------------------------------------------------
# C2: 1 M x 1536, 1024 queries, progressive top-10
# ---------------------------------------------------------------------------------------------------------------
def _check_topk_against_torch(hq, db, emb, q, ids, scores, mask, sample, k):
    """ids / scores returned for the sampled queries == top-k of the exact fp64 cosine over the filter's survivors."""
    from hilbert_quantization_b200.search import unpack_mask
    N = emb.shape[0]
    alive = torch.from_numpy(unpack_mask(mask[sample], N)).cuda()
    qs = q[sample].double()
    cos = torch.empty((len(sample), N), dtype=torch.float64, device="cuda")
    step = 1 << 18
    for s in range(0, N, step):
        c = emb[s:s + step].double()
        cos[:, s:s + step] = (qs @ c.T) / (qs.norm(dim=1)[:, None] * c.norm(dim=1)[None, :])
    sc = (cos + 1.0) * 0.5
    sc[~alive] = -1.0
    for r, j in enumerate(sample):
        got_i, got_s = ids[j].cpu().numpy(), scores[j].double().cpu().numpy()
        n_alive = int(alive[r].sum())
        kk = min(k, n_alive)
        assert (got_i[:kk] >= 0).all() and (got_i[kk:] == -1).all()
        assert alive[r][torch.from_numpy(got_i[:kk]).cuda()].all()                       # only survivors are returned
        assert np.abs(got_s[:kk] - sc[r][torch.from_numpy(got_i[:kk]).cuda()].cpu().numpy()).max() < 5e-7
        assert (np.diff(got_s[:kk]) <= 0).all()                                          # descending
        best = torch.topk(sc[r], kk).values.cpu().numpy()
        assert np.abs(got_s[:kk] - best).max() < 5e-7                                    # and the best there are


def test_c2_search_1m_rows_1024_queries(hq):
    _need_gb(60)
    from hilbert_quantization_b200.search import FilterTrace, rag_ratio, unpack_mask
    N, D, Q, k = 1_000_000, 1536, 1024, 10
    emb = _unit_rows(N, D, 1234)
    q = _queries(emb, Q, 4321)
    q[3] = emb[3]                                                    # an exact copy: must come back first with score 1
    db = hq.EmbeddingDatabase(emb)
    assert db.fast_filter_ok and db.tc_packed is not None
    tr = FilterTrace([], [], [])
    ids, scores, mask = hq.search_batch(db, q, k, return_mask=True, trace=tr)
    L = db.num_levels
    # rag/search/engine.py:279: max(1, int(len * ratio)) caps every level, thresholds only ever remove rows
    prev = torch.full((Q,), N, dtype=torch.int64, device="cuda")
    for l in range(L):
        n_pass, n_out = tr.n_pass[l].to(torch.int64), tr.n_out[l].to(torch.int64)
        cap = torch.clamp((prev.double() * rag_ratio(l)).floor().to(torch.int64), min=1)
        assert torch.equal(tr.n_alive[l].to(torch.int64), prev) and (n_pass <= prev).all()
        assert torch.equal(n_out, torch.minimum(n_pass, cap))
        prev = n_out
    pop = torch.from_numpy(unpack_mask(mask[:64], N).sum(axis=1)).cuda()
    assert torch.equal(pop, prev[:64])                               # the mask holds exactly the last level's survivors
    assert ids[3, 0].item() == 3 and abs(scores[3, 0].item() - 1.0) < 5e-7
    sample = [0, 1, 3, 255, 511, 512, 700, 1023]
    _check_topk_against_torch(hq, db, emb, q, ids, scores, mask, sample, k)
    # the same queries through the exact filter + fp32 rerank (the path that is value-by-value pinned to the oracle at
    # small sizes): survivor sets may differ only by rows within 2e-6 of a threshold / cut score, and the queries
    # whose sets agree return identical ids
    sub = torch.tensor(sample, device="cuda")
    i_e, s_e, m_e = hq.search_batch(db, q[sub], k, return_mask=True, filter_impl="exact", rerank="f32")
    a_f, a_e = unpack_mask(mask[sub], N), unpack_mask(m_e, N)
    diff = (a_f != a_e).sum(axis=1)
    # expected: ~N * 1.8 rows per unit of level score * 4e-7 = under one row per query and level
    assert diff.max() <= 16, diff
    for r in range(len(sample)):
        if diff[r] == 0:
            assert torch.equal(ids[sample[r]], i_e[r])
            assert (scores[sample[r]] - s_e[r]).abs().max().item() < 5e-7
        else:                                                        # a borderline row is almost never one of the best ten
            assert len(set(ids[sample[r]].tolist()) & set(i_e[r].tolist())) >= k - 1


# ---------------------------------------------------------------------------------------------------------------
# C5: one shard of 100 M x 768 over 8 GPUs = 12.5 M rows, 4096-query batch
# ---------------------------------------------------------------------------------------------------------------
def test_c5_shard_12m_rows_4096_queries(hq):
    _need_gb(140)
    N, D, Q, k = 12_500_000, 768, 4096, 10
    emb = _unit_rows(N, D, 99)
    q = _queries(emb, Q, 7)
    q[5] = emb[N - 1]                                                # the shard's last row
    db = hq.EmbeddingDatabase(emb, id_base=3 * N)                    # rank 3 of 8: global ids
    ids, scores, mask = hq.search_batch(db, q, k, return_mask=True)
    assert ids[5, 0].item() == 3 * N + N - 1 and abs(scores[5, 0].item() - 1.0) < 5e-7
    assert ((ids == -1) | ((ids >= 3 * N) & (ids < 4 * N))).all()
    sample = [0, 5, 2047, 2048, 4095]
    _check_topk_against_torch(hq, db, emb, q, ids - 3 * N * (ids >= 0), scores, mask, sample, k)
