"""Generate tests/golden/precomputed.npz from the REAL reference (authoring container only):
PrecomputedHilbertIndexer.create_precomputed_index (core/precomputed_hilbert_index.py:65-212)."""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference, reference_path  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert load_reference() is not None
    from hilbert_quantization.core.precomputed_hilbert_index import PrecomputedHilbertIndexer
    rng = np.random.default_rng(20261020)
    out = {}
    for n in (8, 16, 64, 128):
        img = rng.standard_normal((n, n)).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            idx = PrecomputedHilbertIndexer().create_precomputed_index(img, f"m{n}")
        out[f"n{n}_image"] = img
        out[f"n{n}_levels"] = np.array([(l.grid_size, l.square_size, l.num_squares) for l in idx.levels], dtype=np.int64)
        for i, l in enumerate(idx.levels):
            out[f"n{n}_avg{i}"] = l.averages
            out[f"n{n}_xy{i}"] = np.array(l.square_coordinates, dtype=np.int64).reshape(-1, 2)
    np.savez_compressed(os.path.join(OUT, "precomputed.npz"), **out)
    print("precomputed.npz", os.path.getsize(os.path.join(OUT, "precomputed.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
