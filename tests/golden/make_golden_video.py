"""Generate tests/golden/video_order.npz from the REAL reference (authoring container only): the video-path
hierarchical similarity, greedy frame ordering and insertion position of core/video_storage.py, called as
unbound methods (they only touch `self` for each other and `_current_metadata`)."""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class Meta:
    """stand-in for VideoFrameMetadata (identity equality, like the dataclass behaves for distinct frames)"""

    def __init__(self, hierarchical_indices, frame_index, tag):
        self.hierarchical_indices, self.frame_index, self.tag = hierarchical_indices, frame_index, tag


def main():
    assert load_reference() is not None
    from hilbert_quantization.core.video_storage import VideoModelStorage as V
    sim = lambda q, c: V._calculate_hierarchical_similarity(None, q, c)     # noqa: E731
    rng = np.random.default_rng(20261021)
    out = {}
    for S, F, dt in ((64, 40, np.float32), (1024, 25, np.float32), (32, 30, np.float64)):
        base = rng.standard_normal(S)
        frames = [(base * rng.uniform(0.2, 1.0) + rng.standard_normal(S) * rng.uniform(0.1, 1.5)).astype(dt) for _ in range(F)]
        frames[7] = frames[3].copy()                          # exact duplicate
        frames[11] = np.full(S, 0.5, dtype=dt)                # zero variance
        q = (base + 0.3 * rng.standard_normal(S)).astype(dt)
        tag = f"S{S}_{np.dtype(dt).name}"
        out[f"{tag}_frames"] = np.stack(frames)
        out[f"{tag}_query"] = q
        out[f"{tag}_sims"] = np.array([sim(q, f) for f in frames])
        out[f"{tag}_pair_sims"] = np.array([[sim(a, b) for b in frames[:12]] for a in frames[:12]])
        metas = [Meta(f, i, i) for i, f in enumerate(frames)]
        self = SimpleNamespace(_calculate_hierarchical_similarity=sim)
        ordered = V._sort_frames_by_hierarchical_indices(self, metas)
        out[f"{tag}_order"] = np.array([m.tag for m in ordered], dtype=np.int64)
        self2 = SimpleNamespace(_calculate_hierarchical_similarity=sim, _current_metadata=metas[:15])
        out[f"{tag}_insert_pos"] = np.array([V._find_optimal_insertion_position(self2, q),
                                             V._find_optimal_insertion_position(self2, frames[20])], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "video_order.npz"), **out)
    print("video_order.npz", os.path.getsize(os.path.join(OUT, "video_order.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
