"""GPU: row f2 of SURVEY 8(f) -- enhanced frames as the device-resident database format (explicit original height),
EmbeddingFrame / QuantizedModel batch containers (hilbert_quantization_b200/frames.py)."""
import numpy as np
import pytest
import torch

from oracle import hilbert_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hq():
    import hilbert_quantization_b200 as m
    return m


@pytest.mark.parametrize("D,n", [(1536, 64), (768, 32), (1000, 32)])
def test_frame_batch_equals_the_reference_layout(hq, D, n):
    """frames = map_to_2d grid + one zero-padded index row per level (rag/.../hierarchical_index_generator.py:344-385);
    extraction at the EXPLICIT height returns grid and trimmed rows like :387-441 (oracle c_extract_rows)."""
    rng = np.random.default_rng(D)
    emb = rng.standard_normal((37, D)).astype(np.float32)
    emb[5, D // 2:] = 0.0                                     # trailing zeros inside the grid: the height heuristic's trap
    batch = hq.EmbeddingFrameBatch.from_embeddings(emb, n)
    assert batch.original_height == n and batch.hilbert_dimensions == (n, n) and len(batch) == 37
    host = batch.frames.cpu().numpy()
    for i in (0, 5, 36):
        grid = O.map_to_2d(emb[i], (n, n))
        rows = O.index_c_rows(grid)
        assert np.array_equal(host[i, :n], grid)
        for l, r in enumerate(rows):
            assert np.abs(host[i, n + l, : len(r)] - r).max() <= 3e-7 and not host[i, n + l, len(r):].any()
    frames = batch.to_frames(compression_quality=0.5, first_frame_number=7)
    assert frames[3].frame_number == 10 and frames[3].original_embedding_dimensions == D
    for i in (0, 5):
        g, want_rows = O.c_extract_rows(host[i], n)
        assert np.array_equal(frames[i].embedding_data[:n], g) and len(frames[i].hierarchical_indices) == len(want_rows)
        for a, b in zip(frames[i].hierarchical_indices, want_rows):
            assert np.array_equal(a, b)
    # round trip: embeddings back out of the frames (bit exact), index rows sliced out of the frames
    assert torch.equal(batch.embeddings().cpu(), torch.from_numpy(emb))
    _, idx = hq.map_and_index(torch.from_numpy(emb).cuda(), n, variant="C", layout="compact", want_grid=False)
    assert torch.equal(batch.index_rows(), idx)


def test_database_from_stored_frames_searches_like_the_embedding_database(hq):
    rng = np.random.default_rng(3)
    N, D, n = 20000, 1536, 64
    emb = rng.standard_normal((N, D)).astype(np.float32)
    qs = rng.standard_normal((40, D)).astype(np.float32)
    qs[:10] = emb[:10] + 0.05 * rng.standard_normal((10, D)).astype(np.float32)
    batch = hq.EmbeddingFrameBatch.from_embeddings(emb, n)
    # the stored form: host objects with the reference's fields, loaded back without looking at pixel values for the height
    stored = batch.to_frames()[:64]
    again = hq.EmbeddingFrameBatch.from_frames(stored)
    assert again.original_height == n and torch.equal(again.frames, batch.frames[:64])
    db_f = batch.database()
    db_e = hq.EmbeddingDatabase(emb, n=n)
    assert torch.equal(db_f.idx, db_e.idx) and torch.equal(db_f.emb, db_e.emb) and torch.equal(db_f.lens, db_e.lens)
    i_f, s_f = hq.search_batch(db_f, qs, 10)
    i_e, s_e = hq.search_batch(db_e, qs, 10)
    assert torch.equal(i_f, i_e) and torch.equal(s_f, s_e)
    with pytest.raises(ValueError):
        hq.EmbeddingFrameBatch(batch.frames[:, :-1], n, D)             # a missing index row is an error, not a guess


def test_quantized_model_batch_is_a_device_resident_candidate_pool(hq):
    """core path (models.py:55-79, core/search_engine.py:302-388): the same results from a QuantizedModelBatch as from the
    list of models it wraps; results refer to the original objects."""
    class Model:
        def __init__(self, idx, name):
            self.hierarchical_indices, self.name = idx, name
    rng = np.random.default_rng(9)
    pool = [Model(rng.standard_normal(64 if i % 3 else 32).astype(np.float32), f"m{i}") for i in range(300)]
    q = pool[17].hierarchical_indices + 0.01 * rng.standard_normal(64).astype(np.float32)
    eng = hq.ProgressiveSimilaritySearchEngine(similarity_threshold=0.1, max_candidates_per_level=50)
    batch = hq.QuantizedModelBatch(pool)
    assert len(batch) == 300 and batch[17] is pool[17]
    a = eng.progressive_search(q, pool, 10)
    b = eng.progressive_search(q, batch, 10)
    assert [r.model.name for r in a] == [r.model.name for r in b] and a[0].model is pool[17]
    assert [r.similarity_score for r in a] == [r.similarity_score for r in b]
    c, d = eng.brute_force_search(q, pool, 5), eng.brute_force_search(q, batch, 5)
    assert [r.model.name for r in c] == [r.model.name for r in d]
