"""Generate tests/golden/embed_extract.npz from the REAL reference (authoring container only): row a9 --
HierarchicalIndexGeneratorImpl.embed_indices_in_image / extract_indices_from_image (core/index_generator.py:221-290)
and HierarchicalIndexGenerator.embed_multi_level_indices / extract_indices_from_image with an explicit
original_height (rag/embedding_generation/hierarchical_index_generator.py:344-441)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases(rng):
    """(name, image, indices) for the single-row embed: short / exact / too long index vectors, trailing zeros,
    an all-zero vector, a zero in the middle, float64 and float32 images, a 1-row image."""
    out = []
    for n, dt in ((8, np.float32), (16, np.float64), (32, np.float32)):
        img = rng.standard_normal((n, n)).astype(dt)
        out.append((f"short_{n}", img, rng.standard_normal(n // 2)))
        out.append((f"exact_{n}", img, rng.standard_normal(n).astype(np.float32)))
        out.append((f"long_{n}", img, rng.standard_normal(n + 5)))
        z = rng.standard_normal(n)
        z[-3:] = 0.0
        out.append((f"trailing_zeros_{n}", img, z))
        m = rng.standard_normal(n)
        m[2] = 0.0
        out.append((f"middle_zero_{n}", img, m))
        out.append((f"all_zero_{n}", img, np.zeros(n)))
    out.append(("one_row", rng.standard_normal((1, 8)).astype(np.float32), rng.standard_normal(4)))
    return out


def main():
    assert load_reference() is not None
    from hilbert_quantization.core.index_generator import HierarchicalIndexGeneratorImpl
    from hilbert_quantization.rag.embedding_generation.hierarchical_index_generator import HierarchicalIndexGenerator
    rng = np.random.default_rng(20261019)
    gen, rgen = HierarchicalIndexGeneratorImpl(), HierarchicalIndexGenerator()
    out, names = {}, []
    for name, img, idx in cases(rng):
        enhanced = gen.embed_indices_in_image(img, idx)
        back_img, back_idx = gen.extract_indices_from_image(enhanced)
        names.append(name)
        out[f"a_{name}_image"], out[f"a_{name}_indices"] = img, idx
        out[f"a_{name}_enhanced"], out[f"a_{name}_back_image"], out[f"a_{name}_back_indices"] = enhanced, back_img, back_idx
    out["a_names"] = np.array(names)
    mnames = []
    for n in (8, 32, 64):
        img = rng.standard_normal((n, n)).astype(np.float32)
        rows = [rng.standard_normal(w).astype(np.float32) for w in (n, n // 4, 4)]
        rows[1][-2:] = 0.0                       # trailing zeros are stripped on extraction
        rows.append(np.zeros(3, dtype=np.float32))   # an all-zero row keeps one element
        rows.append(rng.standard_normal(n + 7).astype(np.float32))   # too long: truncated to the width
        enhanced = rgen.embed_multi_level_indices(img, rows)
        back_img, back_rows = rgen.extract_indices_from_image(enhanced, original_height=n)
        name = f"c{n}"
        mnames.append(name)
        out[f"c_{name}_image"], out[f"c_{name}_enhanced"], out[f"c_{name}_back_image"] = img, enhanced, back_img
        out[f"c_{name}_nrows"] = np.array([len(rows), len(back_rows)])
        for i, r in enumerate(rows):
            out[f"c_{name}_row{i}"] = r
        for i, r in enumerate(back_rows):
            out[f"c_{name}_back_row{i}"] = r
    out["c_names"] = np.array(mnames)
    np.savez_compressed(os.path.join(OUT, "embed_extract.npz"), **out)
    print("embed_extract.npz", os.path.getsize(os.path.join(OUT, "embed_extract.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
