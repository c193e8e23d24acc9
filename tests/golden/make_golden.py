"""Generate tests/golden/*.npz by running the REAL reference (not the oracle).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures freeze reference outputs (and the inputs that produced them) so
that the oracle and the CUDA path can be checked on the GPU box, where the
reference is absent.  Each array is named after the reference symbol it came
from.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference, make_quantized_model, rag_filter_with_explicit_rows  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = load_reference()
    assert ref is not None, "reference not reachable"
    rng = np.random.default_rng(20261018)
    mapper = ref.HilbertCurveMapper()

    # ---- coordinates (core/hilbert_mapper.py:17) ----
    coords = {}
    for n in (1, 2, 4, 8, 16, 32, 64, 128):
        c = np.array(mapper.generate_hilbert_coordinates(n), dtype=np.int32)
        coords[f"xy_n{n}"] = c
    # known-answer vectors quoted by the reference's own tests
    coords["test_hilbert_mapper_py_23"] = np.array([(0, 0), (0, 1), (1, 1), (1, 0)], dtype=np.int32)
    coords["test_hilbert_mapper_py_43_prefix4"] = np.array([(0, 0), (1, 0), (1, 1), (0, 1)], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "coords.npz"), **coords)

    # ---- map / unmap (core/hilbert_mapper.py:115,176) ----
    maps = {}
    for n, D, dt in ((2, 3, np.float32), (4, 16, np.int32), (8, 50, np.int64), (16, 200, np.float64),
                     (32, 768, np.float32), (32, 1024, np.float32), (64, 1536, np.float32)):
        p = (rng.standard_normal(D) * 100).astype(dt)
        img = mapper.map_to_2d(p, (n, n))
        key = f"n{n}_D{D}_{np.dtype(dt).name}"
        maps[f"in_{key}"] = p
        maps[f"map_to_2d_{key}"] = img
        maps[f"map_from_2d_{key}"] = mapper.map_from_2d(img)
    maps["golden_from_in"] = np.array([[1, 4], [2, 3]], dtype=np.float64)        # tests/test_hilbert_mapper.py:173-182
    maps["golden_from_out"] = mapper.map_from_2d(maps["golden_from_in"])
    maps["golden_rag_from_in"] = np.array([[1, 2], [3, 4]], dtype=np.float64)    # tests/test_rag_hilbert_mapper.py:691-701
    maps["golden_rag_from_out"] = ref.HilbertCurveMapperImpl(None).map_from_2d(maps["golden_rag_from_in"])
    np.savez_compressed(os.path.join(OUT, "maps.npz"), **maps)

    # ---- index variants A / B / C + embed + uint8 ----
    idx = {}
    agen = ref.HierarchicalIndexGeneratorImpl(ref.QuantizationConfig(use_streaming_optimization=False))
    sgen = ref.StreamingHilbertIndexGenerator()
    cgen = ref.HierarchicalIndexGenerator()
    comp = ref.MPEGAICompressorImpl()
    for n, D in ((8, 64), (16, 200), (32, 768), (32, 1024), (64, 1536), (64, 4096), (128, 16384)):
        p = rng.standard_normal(D).astype(np.float32)
        img = mapper.map_to_2d(p, (n, n))
        key = f"n{n}_D{D}"
        idx[f"in_{key}"] = p
        idx[f"A_{key}"] = agen._generate_traditional_indices(img, n)
        idx[f"B_{key}"] = sgen.generate_optimized_indices(img, n)
        idx[f"C_{key}"] = cgen.generate_multi_level_indices(img)
        enh = agen.embed_indices_in_image(img, idx[f"B_{key}"])          # pipeline.py:133-140 (B into float32 row)
        idx[f"enhB_{key}"] = enh
        idx[f"u8_{key}"] = comp._normalize_for_compression(enh)
        idx[f"deq_{key}"] = comp._denormalize_from_compression(idx[f"u8_{key}"])
    idx["A_S1024_n128"] = agen._generate_traditional_indices(mapper.map_to_2d(idx["in_n128_D16384"], (128, 128)), 1024)
    idx["B_during_mapping_in"] = rng.standard_normal(1000).astype(np.float32)
    idx["B_during_mapping_out"] = sgen.generate_indices_during_mapping(idx["B_during_mapping_in"], (32, 32), 32)[1]
    idx["A_ramp_means"] = np.array(agen.calculate_spatial_averages(np.arange(16, dtype=np.float32).reshape(4, 4), 2))
    np.savez_compressed(os.path.join(OUT, "index.npz"), **idx)

    # ---- RAG progressive search (rag/search/engine.py) ----
    rag = {}
    for tag, n, D, N in (("a", 16, 256, 240), ("b", 32, 768, 160), ("c", 64, 1536, 96)):
        db = rng.standard_normal((N, D)).astype(np.float32)
        db /= np.linalg.norm(db, axis=1, keepdims=True)
        db[9] = db[4]
        qs = []
        for j in range(4):
            if j % 2 == 0:
                q = db[17 + j] + 0.1 * rng.standard_normal(D).astype(np.float32)
            else:
                q = rng.standard_normal(D).astype(np.float32)
            qs.append((q / np.linalg.norm(q)).astype(np.float32))
        qs = np.stack(qs)
        frames = [cgen.generate_multi_level_indices(mapper.map_to_2d(v, (n, n))) for v in db]
        rag[f"{tag}_db"] = db
        rag[f"{tag}_queries"] = qs
        rag[f"{tag}_n"] = np.array(n)
        for j, q in enumerate(qs):
            qf = cgen.generate_multi_level_indices(mapper.map_to_2d(q, (n, n)))
            surv, eng = rag_filter_with_explicit_rows(ref, qf, frames, n)
            rag[f"{tag}_survivors_q{j}"] = np.array(surv, dtype=np.int64)
            ids = sorted(surv)
            sc = [eng._calculate_embedding_cosine_similarity(qf[:n], frames[i][:n]) for i in ids]
            pairs = sorted(zip(ids, sc), key=lambda t: t[1], reverse=True)[:10]
            rag[f"{tag}_topk_ids_q{j}"] = np.array([p[0] for p in pairs], dtype=np.int64)
            rag[f"{tag}_topk_scores_q{j}"] = np.array([p[1] for p in pairs], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "rag_search.npz"), **rag)

    # ---- core progressive search (core/search_engine.py) ----
    core = {}
    ceng = ref.ProgressiveSimilaritySearchEngine(similarity_threshold=0.1, max_candidates_per_level=20)
    for S, N in ((32, 150), (64, 120)):
        cands = rng.standard_normal((N, S))
        cands[5] = 1.25
        cands[6] = cands[2]
        q = cands[2] + 0.05 * rng.standard_normal(S)
        models = [make_quantized_model(ref, cands[i].copy(), f"m{i}") for i in range(N)]
        res = ceng.progressive_search(q, models, 10)
        core[f"S{S}_cands"] = cands
        core[f"S{S}_query"] = q
        core[f"S{S}_ids"] = np.array([int(r.model.metadata.model_name[1:]) for r in res], dtype=np.int64)
        core[f"S{S}_scores"] = np.array([r.similarity_score for r in res])
        nl = len(ceng._parse_index_structure(q, S))
        core[f"S{S}_level_sims"] = np.array([[ceng.compare_indices_at_level(q, cands[i], l) for l in range(nl)]
                                             for i in range(N)])
    np.savez_compressed(os.path.join(OUT, "core_search.npz"), **core)

    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
