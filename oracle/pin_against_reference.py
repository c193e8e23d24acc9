"""TEST INFRASTRUCTURE ONLY -- pin oracle/hilbert_oracle.py against the real reference.

Run in the authoring container (needs /root/reference):

    python -m oracle.pin_against_reference

Every check compares a function of the NumPy restatement with the reference's
own class on seeded inputs.  Exits non-zero on the first mismatch.  The last
recorded run is summarised in DESIGN.md.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from oracle import hilbert_oracle as O
from oracle.reference_loader import load_reference, make_quantized_model, rag_filter_with_explicit_rows

FAILS = []


def check(name, ok, detail=""):
    print(("PASS " if ok else "FAIL ") + name + (f"  [{detail}]" if detail else ""))
    if not ok:
        FAILS.append(name)


def main():
    ref = load_reference()
    if ref is None:
        print("reference not reachable; nothing pinned")
        return 2
    rng = np.random.default_rng(7)
    mapper = ref.HilbertCurveMapper()
    rmapper = ref.HilbertCurveMapperImpl(None)

    # --- a1/a2 coordinates, bit-exact ---
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        want = mapper.generate_hilbert_coordinates(n)
        got = O.hilbert_coordinates_list(n)
        check(f"d2xy n={n}", got == want)
        if n <= 64:
            check(f"d2xy(rag) n={n}", got == rmapper.generate_hilbert_coordinates(n))
        xs = np.array([c[0] for c in want]); ys = np.array([c[1] for c in want])
        d = O.xy2d(n, xs, ys)
        want_d = [mapper._xy_to_hilbert_index(int(x), int(y), n) for x, y in want[: min(len(want), 4096)]]
        check(f"xy2d n={n}", np.array_equal(d, np.arange(n * n)) and list(d[: len(want_d)]) == want_d)

    # --- a3/a4 map/unmap, bit-exact, several dtypes and fills ---
    for n, D, dt in ((2, 3, np.float32), (4, 16, np.int32), (8, 50, np.int64), (16, 200, np.float64),
                     (32, 768, np.float32), (32, 1024, np.float32), (64, 1536, np.float32), (64, 4096, np.float32)):
        p = (rng.standard_normal(D) * 100).astype(dt)
        want = mapper.map_to_2d(p, (n, n))
        got = O.map_to_2d(p, (n, n))
        check(f"map_to_2d n={n} D={D} {np.dtype(dt).name}", got.dtype == want.dtype and np.array_equal(got, want))
        check(f"map_to_2d(rag) n={n} D={D}", np.array_equal(rmapper.map_to_2d(p, (n, n)), got))
        back = mapper.map_from_2d(want)
        check(f"map_from_2d n={n}", np.array_equal(O.map_from_2d(want), back) and back.dtype == want.dtype)
    P = rng.standard_normal((5, 100)).astype(np.float32)
    check("map_to_2d_batch", all(np.array_equal(O.map_to_2d_batch(P, 16)[i], mapper.map_to_2d(P[i], (16, 16))) for i in range(5)))

    # --- a5 dimensions ---
    calc = ref.PowerOf4DimensionCalculator()
    for c in (1, 3, 4, 5, 16, 17, 64, 65, 768, 1024, 1025, 1536, 4096, 5000, 16384, 16385, 70000, 300000, 16777216):
        check(f"dims {c}", O.optimal_dimensions(c) == calc.calculate_optimal_dimensions(c))
    gen_dims = ref.EmbeddingGeneratorImpl.calculate_optimal_dimensions
    for c in (1, 2, 5, 64, 384, 768, 1024, 1536, 4096, 4097):
        check(f"rag dims {c}", O.rag_optimal_dimensions(c) == gen_dims(None, c))

    # --- a6 index B, bit-exact float64 ---
    sgen = ref.StreamingHilbertIndexGenerator()
    for n, S in ((2, 2), (4, 4), (4, 7), (8, 8), (16, 16), (32, 32), (64, 64), (64, 100), (128, 128), (32, 5)):
        img = rng.standard_normal((n, n)).astype(np.float32)
        want = sgen.generate_optimized_indices(img, S)
        got = O.index_b(img, S)
        check(f"index B n={n} S={S}", got.dtype == want.dtype and np.array_equal(got, want))
    # padded image (fill 0.375) and the during-mapping flavour (only real values fed)
    p = rng.standard_normal(1536).astype(np.float32)
    img = mapper.map_to_2d(p, (64, 64))
    check("index B padded 64", np.array_equal(O.index_b(img, 64), sgen.generate_optimized_indices(img, 64)))
    _, idx_dm, stats = sgen.generate_indices_during_mapping(p[:1000], (32, 32), 32)
    check("index B during-mapping", np.array_equal(O.index_b_from_values(p[:1000], 32), idx_dm)
          and stats["total_values_processed"] == 1000)

    # --- a7 index A ---
    agen = ref.HierarchicalIndexGeneratorImpl(ref.QuantizationConfig(use_streaming_optimization=False)) \
        if "use_streaming_optimization" in ref.QuantizationConfig.__dataclass_fields__ else ref.HierarchicalIndexGeneratorImpl()
    for S in (1, 2, 3, 5, 10, 16, 32, 64, 100, 256, 1000, 1024, 4096):
        check(f"A alloc S={S}", O.a_level_allocation(S) == agen.calculate_level_allocation(S))
    for n, S in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (64, 20), (128, 128), (256, 256), (256, 1024), (32, 1024)):
        img = rng.standard_normal((n, n)).astype(np.float32)
        want = agen._generate_traditional_indices(img, S)
        got = O.index_a(img, S)
        err = np.abs(got.astype(np.float64) - want).max()
        check(f"index A n={n} S={S}", got.shape == want.shape and got.dtype == want.dtype and err <= 2e-7, f"max|d|={err:.2e}")
    ramp = np.arange(16, dtype=np.float32).reshape(4, 4)
    check("A golden ramp (tests/test_index_generator.py:115-139)",
          np.allclose(O.a_spatial_averages(ramp, 2), agen.calculate_spatial_averages(ramp, 2)))
    img = rng.standard_normal((32, 32)).astype(np.float32)
    idx = agen._generate_traditional_indices(img, 32)
    enh = agen.embed_indices_in_image(img, idx)
    check("embed row", np.array_equal(O.embed_index_row(img, idx), enh))
    a, b = agen.extract_indices_from_image(enh)
    a2, b2 = O.extract_index_row(enh)
    check("extract row", np.array_equal(a, a2) and np.array_equal(b, b2))

    # --- a8 index C ---
    cgen = ref.HierarchicalIndexGenerator()
    for g in (1, 2, 4, 8, 16, 32, 64):
        check(f"C section order g={g}", O.c_section_order(g) == cgen._generate_hilbert_coordinates(g))
    for w in (2, 4, 8, 16, 32, 64, 100, 128, 256, 1024, 4096):
        info = cgen.calculate_optimal_granularity((w, w))
        check(f"C levels W={w}", O.c_granularity_levels(w) == info["granularity_levels"])
    for n, D in ((4, 16), (8, 64), (16, 256), (32, 768), (32, 1024), (64, 1536), (64, 4096), (128, 16384)):
        p = rng.standard_normal(D).astype(np.float32)
        img = mapper.map_to_2d(p, (n, n))
        want = cgen.generate_multi_level_indices(img)
        got = O.index_c(img)
        err = np.abs(got.astype(np.float64) - want).max()
        check(f"index C n={n} D={D}", got.shape == want.shape and got.dtype == want.dtype and err <= 2e-7, f"max|d|={err:.2e}")
        compact = O.index_c_batch_compact(img[None])[0]
        o, ok = 0, True
        for i, g in enumerate(O.c_granularity_levels(n)):
            ok &= np.abs(compact[o:o + g * g] - want[n + i, : g * g]).max() <= 2e-7
            o += g * g
        check(f"index C compact n={n}", ok)
        _, rows_w = cgen.extract_indices_from_image(want, original_height=n)
        _, rows_g = O.c_extract_rows(want, n)
        check(f"C extract n={n}", len(rows_w) == len(rows_g) and all(np.array_equal(a, b) for a, b in zip(rows_w, rows_g)))
    quad = np.zeros((4, 4), dtype=np.float32); quad[:2, :2] = 2; quad[:2, 2:] = 3; quad[2:, 2:] = 4; quad[2:, :2] = 5
    check("C golden quadrants (tests/test_hierarchical_index_generator.py:121-138)",
          np.array_equal(O.c_index_row(quad, 2), np.array([2, 3, 4, 5], dtype=np.float32)))
    rect = rng.standard_normal((6, 10)).astype(np.float32)
    err = np.abs(O.index_c(rect).astype(np.float64) - cgen.generate_multi_level_indices(rect)).max()
    check("index C non-square", err <= 2e-7, f"{err:.2e}")

    # --- a10 uint8 quantise, bit-exact ---
    comp = ref.MPEGAICompressorImpl()
    for shape in ((33, 32), (65, 64), (8, 8)):
        img = (rng.standard_normal(shape) * 3).astype(np.float32)
        want = comp._normalize_for_compression(img)
        got, mn, mx = O.normalize_u8(img)
        check(f"u8 normalise {shape}", np.array_equal(got, want) and got.dtype == want.dtype)
        back = comp._denormalize_from_compression(want)
        check(f"u8 denormalise {shape}", np.array_equal(O.denormalize_u8(want, mn, mx), back))
    const = np.full((4, 4), 2.5, dtype=np.float32)
    check("u8 constant", np.array_equal(O.normalize_u8(const)[0], ref.MPEGAICompressorImpl()._normalize_for_compression(const)))

    # --- a12/a13/a15 RAG filter + rerank + top-k ---
    for n, D, N, seed in ((32, 768, 400, 1), (64, 1536, 300, 2), (32, 1024, 500, 3), (16, 256, 200, 4)):
        r = np.random.default_rng(seed)
        db = r.standard_normal((N, D)).astype(np.float32)
        db /= np.linalg.norm(db, axis=1, keepdims=True)
        db[7] = db[3]                                    # exact duplicates -> tie handling
        q = db[11] + 0.1 * r.standard_normal(D).astype(np.float32)
        q = (q / np.linalg.norm(q)).astype(np.float32)
        frames = [cgen.generate_multi_level_indices(mapper.map_to_2d(v, (n, n))) for v in db]
        qf = cgen.generate_multi_level_indices(mapper.map_to_2d(q, (n, n)))
        want, eng = rag_filter_with_explicit_rows(ref, qf, frames, n)
        L = len(O.c_granularity_levels(n))
        c_rows = [np.stack([f[n + l] for f in frames]) for l in range(L)]
        q_rows = [qf[n + l] for l in range(L)]
        got, trace = O.rag_progressive_filter(q_rows, c_rows, return_scores=True)
        same = list(got) == list(want)
        if not same:   # tolerate only borderline (|s-thr|<1e-6) differences
            diff = set(got) ^ set(want)
            border = set()
            for cand, s, thr, cap in trace:
                border |= set(cand[np.abs(s - thr) < 1e-6].tolist())
            same = diff <= border
        check(f"RAG filter n={n} N={N}", same, f"{len(want)} survivors")
        # level-0 scores vs the reference's own comparator
        cand, s, thr, cap = trace[0]
        gen_rows = [cgen.extract_indices_from_image(f, original_height=n)[1] for f in frames[:50]]
        qr = cgen.extract_indices_from_image(qf, original_height=n)[1]
        ref_s = []
        for rws in gen_rows:
            m = min(len(qr[0]), len(rws[0]))
            ref_s.append(eng._compare_single_level_indices(qr[0][:m], rws[0][:m]))
        check(f"RAG level score n={n}", np.abs(np.array(ref_s) - s[:50]).max() < 5e-7)
        # rerank + top-k
        surv = sorted(want)
        ref_scores = [eng._calculate_embedding_cosine_similarity(qf[:n], frames[i][:n]) for i in surv]
        pairs = sorted(zip(surv, ref_scores), key=lambda t: t[1], reverse=True)[:10]
        ids, sc = O.progressive_search(q, db, n, 10, db_rows=c_rows)
        ok = [p[0] for p in pairs] == list(ids) and np.abs(np.array([p[1] for p in pairs]) - sc).max() < 5e-7
        check(f"RAG rerank/top-k n={n}", ok, f"top1={ids[0] if len(ids) else None}")

    # --- a11 core search ---
    ceng = ref.ProgressiveSimilaritySearchEngine(similarity_threshold=0.1, max_candidates_per_level=20)
    for S in (16, 32, 64):
        check(f"core parse S={S}", O.core_parse_levels(S, S) ==
              [(l.grid_size, l.start_index, l.end_index, l.is_offset_sampling) for l in ceng._parse_index_structure(np.zeros(S), S)])
    for S, N, seed in ((32, 150, 5), (64, 120, 6)):
        r = np.random.default_rng(seed)
        cands = r.standard_normal((N, S))
        cands[5] = 1.25          # constant vector special cases
        cands[6] = cands[2]
        q = cands[2] + 0.05 * r.standard_normal(S)
        models = [make_quantized_model(ref, cands[i].copy(), f"m{i}") for i in range(N)]
        sims = O.core_all_level_similarities(q, cands)
        nl = sims.shape[1]
        ref_sims = np.array([[ceng.compare_indices_at_level(q, cands[i], l) for l in range(nl)] for i in range(40)])
        check(f"core level sims S={S}", np.abs(ref_sims - sims[:40]).max() < 1e-12, f"{np.abs(ref_sims - sims[:40]).max():.1e}")
        res = ceng.progressive_search(q, models, 10)
        ids, sc = O.core_progressive_search(q, cands, 10, 0.1, 20)
        want_ids = [int(x.model.metadata.model_name[1:]) for x in res]
        check(f"core progressive S={S}", want_ids == list(ids) and
              np.abs(np.array([x.similarity_score for x in res]) - sc).max() < 1e-12, f"top={want_ids[:3]}")

    # --- a14 comprehensive blend (rag/search/engine.py:516-575), index rows / original rows read explicitly ---
    from unittest.mock import patch
    eng = ref.RAGSearchEngineImpl(ref.RAGConfig())
    gen = ref.HierarchicalIndexGenerator()
    for n, D in ((64, 1536), (32, 768), (16, 200), (64, 4096)):
        r = np.random.default_rng(n + D)

        def frame(v):
            return gen.generate_multi_level_indices(rmapper.map_to_2d(v.astype(np.float32), (n, n))).astype(np.float32)
        qv = r.standard_normal(D)
        cfs = [frame(r.standard_normal(D)) for _ in range(5)] + [frame(qv + 0.1 * r.standard_normal(D))]
        qf = frame(qv)
        with patch.object(eng, "_extract_hierarchical_indices", side_effect=lambda fr: gen.extract_indices_from_image(fr, original_height=n)[1]), \
                patch.object(eng, "_extract_original_embedding", side_effect=lambda fr: fr[:n, :] if fr.ndim == 2 else fr):
            qi = eng._extract_hierarchical_indices(qf)
            want = np.array([eng._calculate_comprehensive_similarity(qf, qi, c, 0) for c in cfs])
            want_sp = np.array([eng._calculate_spatial_locality_similarity(qf, c) for c in cfs])
        got = O.comprehensive_similarity(qf, np.stack(cfs), n)
        check(f"comprehensive blend n={n} D={D}", np.abs(want - got).max() < 2e-7, f"{np.abs(want - got).max():.1e}")
        got_sp = O.spatial_locality_similarity(qf[:n], np.stack(cfs)[:, :n])
        check(f"spatial locality n={n} D={D}", np.abs(want_sp - got_sp).max() < 2e-7, f"{np.abs(want_sp - got_sp).max():.1e}")
    for L in (1, 2, 3, 6):
        check(f"granularity weights L={L}", np.allclose(O.granularity_weights(L), eng._calculate_granularity_weights(L), rtol=0, atol=0))

    # --- f1 PrecomputedHilbertIndexer (core/precomputed_hilbert_index.py:122-212) ---
    import contextlib
    import io
    from hilbert_quantization.core.precomputed_hilbert_index import PrecomputedHilbertIndexer
    for n in (8, 32, 64):
        img = rng.standard_normal((n, n)).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            pidx = PrecomputedHilbertIndexer().create_precomputed_index(img, f"pin{n}")
        lv = O.precomputed_granularity_levels(n)
        check(f"precomputed levels n={n}", lv == [(l.grid_size, l.square_size) for l in pidx.levels])
        ok = all(np.array_equal(O.precomputed_level_averages(img, g, s), l.averages) for (g, s), l in zip(lv, pidx.levels))
        check(f"precomputed averages n={n}", ok)

    # --- f4 video path (core/video_storage.py:763-781,1203-1277,1751-1803; core/video_search.py:215-262,1316-1328) ---
    from types import SimpleNamespace
    from hilbert_quantization.core.video_search import VideoEnhancedSearchEngine as VE
    from hilbert_quantization.core.video_storage import VideoModelStorage as VS
    for S, F, thr in ((64, 30, 0.1), (340, 20, 0.55)):
        base = rng.standard_normal(S)
        frames = [base * rng.uniform(-0.5, 1.0) + rng.standard_normal(S) * rng.uniform(0.05, 1.5) for _ in range(F)]
        frames[7] = frames[3].copy()
        frames[11] = np.full(S, 0.5)
        q = base + 0.3 * rng.standard_normal(S)
        want = np.array([VS._calculate_hierarchical_similarity(None, q, f) for f in frames])
        got = np.array([O.video_hierarchical_similarity(q, f) for f in frames])
        check(f"video pearson similarity S={S}", np.array_equal(want, got))
        metas = [SimpleNamespace(hierarchical_indices=None if i == 4 else f, frame_index=i) for i, f in enumerate(frames)]
        eng_c = ref.ProgressiveSimilaritySearchEngine(thr, 100)
        me = SimpleNamespace(video_storage=SimpleNamespace(_video_index={"v": SimpleNamespace(frame_metadata=metas)}),
                             similarity_threshold=thr, traditional_engine=eng_c)
        me._calculate_hierarchical_similarity = lambda a, b, _s=me: VE._calculate_hierarchical_similarity(_s, a, b)
        res = VE._hierarchical_search(me, SimpleNamespace(hierarchical_indices=q), 8)
        mine = O.video_hierarchical_search(q, [m.hierarchical_indices for m in metas], 8, thr)
        check(f"video hierarchical search S={S}", [r.frame_metadata.frame_index for r in res] == [m[0] for m in mine]
              and np.abs(np.array([r.similarity_score for r in res]) - np.array([m[1] for m in mine])).max() < 1e-12)

    print(f"\n{len(FAILS)} failure(s)")
    return 1 if FAILS else 0


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"elapsed {time.time() - t0:.1f}s")
    sys.exit(rc)
