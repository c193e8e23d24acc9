"""Generate tests/golden/video_search.npz from the REAL reference (authoring container only):
core/video_search.py:215-262 `VideoEnhancedSearchEngine._hierarchical_search`, called as an unbound method over a
stand-in storage object (it only touches `video_storage._video_index`, `similarity_threshold` and
`_calculate_hierarchical_similarity` -> the real core engine's `compare_indices_at_level(q, c, 0)`)."""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = load_reference()
    assert ref is not None
    from hilbert_quantization.core.video_search import VideoEnhancedSearchEngine as E
    rng = np.random.default_rng(20261022)
    out = {}
    for S, F, thr in ((64, 60, 0.1), (340, 45, 0.55), (21, 30, 0.5)):
        base = rng.standard_normal(S)
        frames = [base * rng.uniform(-0.5, 1.0) + rng.standard_normal(S) * rng.uniform(0.05, 1.5) for _ in range(F)]
        frames[7] = frames[3].copy()                          # exact duplicate: a score tie, position order decides
        frames[11] = np.full(S, 0.5)                          # constant frame -> 0.1
        q = base + 0.3 * rng.standard_normal(S)
        missing = [4, 19]                                     # frames without index vectors are skipped (:237)
        metas = [SimpleNamespace(hierarchical_indices=None if i in missing else f, frame_index=i) for i, f in enumerate(frames)]
        videos = {"a": SimpleNamespace(frame_metadata=metas[: F // 2]), "b": SimpleNamespace(frame_metadata=metas[F // 2:])}
        eng = ref.ProgressiveSimilaritySearchEngine(thr, 100)
        self = SimpleNamespace(video_storage=SimpleNamespace(_video_index=videos), similarity_threshold=thr, traditional_engine=eng)
        self._calculate_hierarchical_similarity = lambda a, b, _s=self: E._calculate_hierarchical_similarity(_s, a, b)
        res = E._hierarchical_search(self, SimpleNamespace(hierarchical_indices=q), 12)
        assert res and all(r.search_method == "hierarchical" for r in res)
        tag = f"S{S}"
        out[f"{tag}_frames"] = np.stack(frames)
        out[f"{tag}_missing"] = np.array(missing, dtype=np.int64)
        out[f"{tag}_query"] = q
        out[f"{tag}_threshold"] = np.array(thr)
        out[f"{tag}_ids"] = np.array([r.frame_metadata.frame_index for r in res], dtype=np.int64)
        out[f"{tag}_scores"] = np.array([r.similarity_score for r in res])
        print(tag, len(res), out[f"{tag}_ids"][:6], out[f"{tag}_scores"][:3])
    np.savez_compressed(os.path.join(OUT, "video_search.npz"), **out)
    print("video_search.npz", os.path.getsize(os.path.join(OUT, "video_search.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
