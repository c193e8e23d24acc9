"""CPU: the oracle restatement against fixtures frozen from the real reference
(tests/golden/make_golden.py).  Bit-exact for integer/index/byte work; float
tolerances are stated inline."""
import numpy as np
import pytest

from oracle import hilbert_oracle as O
from conftest import load_golden


def test_coords_bit_exact():
    g = load_golden("coords.npz")
    for n in (1, 2, 4, 8, 16, 32, 64, 128):
        x, y = O.hilbert_coordinates(n)
        want = g[f"xy_n{n}"]
        assert np.array_equal(x, want[:, 0]) and np.array_equal(y, want[:, 1])
        assert np.array_equal(O.xy2d(n, x, y), np.arange(n * n))
    assert O.hilbert_coordinates_list(2) == [tuple(r) for r in g["test_hilbert_mapper_py_23"].tolist()]
    assert O.hilbert_coordinates_list(4)[:4] == [tuple(r) for r in g["test_hilbert_mapper_py_43_prefix4"].tolist()]


def test_map_unmap_bit_exact():
    g = load_golden("maps.npz")
    keys = [k[3:] for k in g.files if k.startswith("in_")]
    assert len(keys) == 7
    for key in keys:
        n = int(key.split("_")[0][1:])
        p = g[f"in_{key}"]
        img = O.map_to_2d(p, (n, n))
        assert img.dtype == g[f"map_to_2d_{key}"].dtype
        assert np.array_equal(img, g[f"map_to_2d_{key}"])
        assert np.array_equal(O.map_from_2d(img), g[f"map_from_2d_{key}"])
        assert np.array_equal(O.map_to_2d_batch(p[None], n)[0], img)
    assert np.array_equal(O.map_from_2d(g["golden_from_in"]), g["golden_from_out"])
    assert list(g["golden_from_out"]) == [1, 2, 3, 4]
    assert list(O.map_from_2d(g["golden_rag_from_in"])) == [1, 3, 4, 2]


@pytest.mark.parametrize("count,dim", [(1, 2), (4, 2), (5, 4), (768, 32), (1024, 32), (1025, 64), (1536, 64),
                                        (4096, 64), (16385, 256), (16777216, 4096), (494032768 // 30, 4096)])
def test_dimensions(count, dim):
    assert O.optimal_dimensions(count) == (dim, dim)


def test_dimension_gate():
    assert not O.efficiency_ok(1536, (64, 64))           # 0.375 < 0.5
    assert O.efficiency_ok(768, (32, 32))
    assert O.rag_optimal_dimensions(1536) == (64, 64)


def _index_keys(g):
    return [k[3:] for k in g.files if k.startswith("in_n")]


def test_index_variants():
    g = load_golden("index.npz")
    for key in _index_keys(g):
        n = int(key.split("_")[0][1:])
        img = O.map_to_2d(g[f"in_{key}"], (n, n))
        # B: float64 strict-order adds -> bit exact
        b = O.index_b(img, n)
        assert b.dtype == np.float64 and np.array_equal(b, g[f"B_{key}"])
        # A / C: float32 block means, reference rounds pairwise in fp32 -> 2e-7 abs on N(0,1) data
        a = O.index_a(img, n)
        assert a.dtype == np.float32 and np.abs(a - g[f"A_{key}"]).max() <= 2e-7
        c = O.index_c(img)
        assert c.shape == g[f"C_{key}"].shape and np.abs(c - g[f"C_{key}"]).max() <= 2e-7
        assert np.array_equal(c[:n], img)
        enh = O.embed_index_row(img, b)
        assert np.array_equal(enh, g[f"enhB_{key}"])
        u8, mn, mx = O.normalize_u8(enh)
        assert np.array_equal(u8, g[f"u8_{key}"])
        assert np.array_equal(O.denormalize_u8(u8, mn, mx), g[f"deq_{key}"])
    img = O.map_to_2d(g["in_n128_D16384"], (128, 128))
    assert np.abs(O.index_a(img, 1024) - g["A_S1024_n128"]).max() <= 2e-7
    assert np.array_equal(O.index_b_from_values(g["B_during_mapping_in"], 32), g["B_during_mapping_out"])
    assert np.allclose(O.a_spatial_averages(np.arange(16, dtype=np.float32).reshape(4, 4), 2), [2.5, 4.5, 10.5, 12.5])
    assert np.allclose(g["A_ramp_means"], [2.5, 4.5, 10.5, 12.5])


def test_index_c_compact_matches_rows():
    g = load_golden("index.npz")
    for key in _index_keys(g):
        n = int(key.split("_")[0][1:])
        want = g[f"C_{key}"]
        compact = O.index_c_batch_compact(want[None, :n])[0]
        o = 0
        for i, gr in enumerate(O.c_granularity_levels(n)):
            assert np.abs(compact[o:o + gr * gr] - want[n + i, : gr * gr]).max() <= 2e-7
            o += gr * gr


def test_rag_search_golden():
    g = load_golden("rag_search.npz")
    for tag in "abc":
        db, qs, n = g[f"{tag}_db"], g[f"{tag}_queries"], int(g[f"{tag}_n"])
        levels = O.c_granularity_levels(n)
        compact = O.index_c_batch_compact(O.map_to_2d_batch(db, n))
        rows, o = [], 0
        for gr in levels:
            r = np.zeros((db.shape[0], n), dtype=np.float32)
            r[:, : gr * gr] = compact[:, o:o + gr * gr]
            rows.append(r)
            o += gr * gr
        for j, q in enumerate(qs):
            qc = O.index_c_batch_compact(O.map_to_2d_batch(q[None], n))[0]
            q_rows, o = [], 0
            for gr in levels:
                r = np.zeros(n, dtype=np.float32)
                r[: gr * gr] = qc[o:o + gr * gr]
                q_rows.append(r)
                o += gr * gr
            surv = O.rag_progressive_filter(q_rows, rows)
            assert list(surv) == list(g[f"{tag}_survivors_q{j}"])
            ids, sc = O.progressive_search(q, db, n, 10, db_rows=rows)
            assert list(ids) == list(g[f"{tag}_topk_ids_q{j}"])
            # reference scores are fp32; oracle accumulates in fp64 -> 5e-7 abs
            assert np.abs(sc - g[f"{tag}_topk_scores_q{j}"]).max() < 5e-7


def test_core_search_golden():
    g = load_golden("core_search.npz")
    for S in (32, 64):
        c, q = g[f"S{S}_cands"], g[f"S{S}_query"]
        sims = O.core_all_level_similarities(q, c)
        assert np.abs(sims - g[f"S{S}_level_sims"]).max() < 1e-12
        ids, sc = O.core_progressive_search(q, c, 10, 0.1, 20)
        assert list(ids) == list(g[f"S{S}_ids"])
        assert np.abs(sc - g[f"S{S}_scores"]).max() < 1e-12


def test_filter_constants():
    assert [round(O.rag_threshold(l), 10) for l in range(5)] == [0.6, 0.5, 0.4, 0.3, 0.3]
    assert [O.rag_ratio(l) for l in range(4)] == [0.3, 0.5, 0.7, 0.7]
    assert O.c_granularity_levels(32) == [4, 2] and O.c_granularity_levels(64) == [8, 4, 2]
    assert O.c_granularity_levels(4096) == [64, 32, 16, 8, 4, 2]
    assert O.a_level_allocation(64) == [(8, 32), (4, 8), (2, 3), (1, 1), (8, 20)]


def test_oracle_comprehensive_blend_matches_reference_golden():
    """a14: oracle restatement of rag/search/engine.py:516-575 against outputs of the reference itself."""
    g = load_golden("rag_blend.npz")
    for tag in ("n64_D1536", "n32_D768", "n16_D200", "n64_D4096"):
        qf, cfs = g[f"{tag}_query_frame"], g[f"{tag}_cand_frames"]
        H = qf.shape[1]
        got = O.comprehensive_similarity(qf, cfs, H)
        assert np.abs(got - g[f"{tag}_comprehensive"]).max() < 2e-7
        assert np.abs(O.spatial_locality_similarity(qf[:H], cfs[:, :H]) - g[f"{tag}_spatial"]).max() < 2e-7
        order = np.argsort(-got, kind="stable")
        assert list(order[:5]) == list(g[f"{tag}_ranked_ids"][:5])


def test_oracle_precomputed_indexer_matches_reference_golden():
    """f1: oracle restatement of core/precomputed_hilbert_index.py:122-212 against the reference's output."""
    g = load_golden("precomputed.npz")
    for n in (8, 16, 64, 128):
        img = g[f"n{n}_image"]
        levels = O.precomputed_granularity_levels(n)
        assert [tuple(r[:2]) for r in g[f"n{n}_levels"]] == levels
        for i, (gs, ss) in enumerate(levels):
            got = O.precomputed_level_averages(img, gs, ss)
            assert got.shape == g[f"n{n}_avg{i}"].shape and np.array_equal(got, g[f"n{n}_avg{i}"])


def test_oracle_video_path_matches_reference_golden():
    """f4: oracle restatement of core/video_storage.py:763-781, :1203-1277, :1751-1803 vs the reference's outputs."""
    g = load_golden("video_order.npz")
    for tag in ("S64_float32", "S1024_float32", "S32_float64"):
        frames, q = list(g[f"{tag}_frames"]), g[f"{tag}_query"]
        sims = np.array([O.video_hierarchical_similarity(q, f) for f in frames])
        assert np.array_equal(sims, g[f"{tag}_sims"])
        assert list(O.video_sort_frames(frames)) == list(g[f"{tag}_order"])
        assert [O.video_insertion_position(q, frames[:15]), O.video_insertion_position(frames[20], frames[:15])] == list(g[f"{tag}_insert_pos"])


def test_oracle_video_hierarchical_search_matches_reference_golden():
    """f4: oracle restatement of core/video_search.py:215-262 (+ :1316-1328) vs VideoEnhancedSearchEngine._hierarchical_search."""
    g = load_golden("video_search.npz")
    for tag in ("S64", "S340", "S21"):
        frames = [None if i in set(g[f"{tag}_missing"].tolist()) else f for i, f in enumerate(g[f"{tag}_frames"])]
        res = O.video_hierarchical_search(g[f"{tag}_query"], frames, 12, float(g[f"{tag}_threshold"]))
        assert [r[0] for r in res] == list(g[f"{tag}_ids"])
        assert np.abs(np.array([r[1] for r in res]) - g[f"{tag}_scores"]).max() < 1e-12
