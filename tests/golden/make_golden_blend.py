"""Generate tests/golden/rag_blend.npz by running the REAL reference (authoring container only):

    python tests/golden/make_golden_blend.py

Comprehensive similarity (rag/search/engine.py:516-575) of query frames against candidate frames,
through RAGSearchEngineImpl._calculate_comprehensive_similarity / calculate_embedding_similarity with the
index rows and the original embedding read explicitly at original_height (the reference's own
height heuristic, :134-162, is patched out exactly like the filter fixtures do).
"""
from __future__ import annotations

import os
import sys
from unittest.mock import patch

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = load_reference()
    assert ref is not None, "reference not reachable"
    rng = np.random.default_rng(20261019)
    eng = ref.RAGSearchEngineImpl(ref.RAGConfig())
    gen = ref.HierarchicalIndexGenerator()
    mapper = ref.HilbertCurveMapperImpl(None)
    out = {}
    for n, D, N in ((64, 1536, 24), (32, 768, 20), (16, 200, 12), (64, 4096, 8)):
        def frame(v):
            return gen.generate_multi_level_indices(mapper.map_to_2d(v.astype(np.float32), (n, n))).astype(np.float32)
        qv = rng.standard_normal(D)
        cands = rng.standard_normal((N, D))
        cands[3] = qv + 0.1 * rng.standard_normal(D)          # a near duplicate of the query
        cands[7] = cands[2]                                     # an exact tie
        qf = frame(qv)
        cfs = [frame(c) for c in cands]
        H = n

        def rows(fr):
            return gen.extract_indices_from_image(fr, original_height=H)[1]

        def orig(fr):
            return fr[:H, :] if fr.ndim == 2 else fr
        with patch.object(eng, "_extract_hierarchical_indices", side_effect=rows), \
                patch.object(eng, "_extract_original_embedding", side_effect=orig):
            qi = eng._extract_hierarchical_indices(qf)
            scores = np.array([eng._calculate_comprehensive_similarity(qf, qi, c, i) for i, c in enumerate(cfs)])
            ranked = eng.calculate_embedding_similarity(qf, {i: c for i, c in enumerate(cfs)})
            spatial = np.array([eng._calculate_spatial_locality_similarity(qf, c) for c in cfs])
        tag = f"n{n}_D{D}"
        out[f"{tag}_query_frame"] = qf
        out[f"{tag}_cand_frames"] = np.stack(cfs)
        out[f"{tag}_comprehensive"] = scores
        out[f"{tag}_spatial"] = spatial
        out[f"{tag}_ranked_ids"] = np.array([r[0] for r in ranked], dtype=np.int64)
        out[f"{tag}_ranked_scores"] = np.array([r[1] for r in ranked])
    np.savez_compressed(os.path.join(OUT, "rag_blend.npz"), **out)
    print("rag_blend.npz", os.path.getsize(os.path.join(OUT, "rag_blend.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
