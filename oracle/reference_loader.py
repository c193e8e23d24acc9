"""TEST INFRASTRUCTURE ONLY -- import the real reference package when it is reachable.

The GPU box never has ``/root/reference``; there the reference is only reachable as the
``baseline/_ref`` install (``baseline/install_reference.sh``), which the CPU-baseline leg of
``bench.py`` times.  Used in the authoring container by ``oracle/pin_against_reference.py``
and ``tests/golden/make_golden*.py``.
"""
from __future__ import annotations

import logging
import os
import sys
import warnings
from types import SimpleNamespace
from unittest.mock import patch


_BASELINE_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_path():
    """$HQ_REFERENCE_PATH -> /root/reference -> baseline/_ref (the install that travels to the GPU box,
    baseline/install_reference.sh)."""
    for p in (os.environ.get("HQ_REFERENCE_PATH"), "/root/reference", _BASELINE_REF):
        if p and os.path.isdir(os.path.join(p, "hilbert_quantization")):
            return p
    return None


def load_reference():
    """Return a namespace of the reference classes used as the ground truth, or None."""
    path = reference_path()
    if path is None:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    warnings.filterwarnings("ignore")
    logging.disable(logging.CRITICAL)
    try:
        from hilbert_quantization.core.hilbert_mapper import HilbertCurveMapper
        from hilbert_quantization.rag.embedding_generation.hilbert_mapper import HilbertCurveMapperImpl
        from hilbert_quantization.core.dimension_calculator import PowerOf4DimensionCalculator
        from hilbert_quantization.core.index_generator import HierarchicalIndexGeneratorImpl
        from hilbert_quantization.core.streaming_index_builder import (
            StreamingHilbertIndexGenerator, StreamingIndexBuilder)
        from hilbert_quantization.rag.embedding_generation.hierarchical_index_generator import (
            HierarchicalIndexGenerator)
        from hilbert_quantization.core.compressor import MPEGAICompressorImpl
        from hilbert_quantization.core.search_engine import ProgressiveSimilaritySearchEngine
        from hilbert_quantization.rag.search.engine import RAGSearchEngineImpl
        from hilbert_quantization.rag.config import RAGConfig
        from hilbert_quantization.rag.embedding_generation.generator import EmbeddingGeneratorImpl
        from hilbert_quantization.models import QuantizedModel, ModelMetadata
        from hilbert_quantization.config import QuantizationConfig
        from hilbert_quantization.exceptions import HilbertQuantizationError
    finally:
        logging.disable(logging.NOTSET)
    return SimpleNamespace(**{k: v for k, v in locals().items() if k[0].isupper()})


def rag_filter_with_explicit_rows(ref, q_frame, cand_frames, H):
    """SURVEY 8c composition: RAGSearchEngineImpl.progressive_hierarchical_search with
    ``_get_all_candidate_embeddings`` patched to the frame list (as the reference's
    tests/test_progressive_filtering.py:167-169 does) and rows fed explicitly
    through HierarchicalIndexGenerator.extract_indices_from_image(frame, original_height=H)."""
    eng = ref.RAGSearchEngineImpl(ref.RAGConfig())
    gen = ref.HierarchicalIndexGenerator()

    def rows(frame):
        if frame.ndim != 2:
            return []
        return gen.extract_indices_from_image(frame, original_height=H)[1]

    with patch.object(eng, "_get_all_candidate_embeddings", return_value=cand_frames), \
            patch.object(eng, "_extract_hierarchical_indices", side_effect=rows):
        return eng.progressive_hierarchical_search(q_frame), eng


def make_quantized_model(ref, indices, name="m"):
    """Like tests/test_search_engine.py:32-47."""
    return ref.QuantizedModel(
        compressed_data=b"x", original_dimensions=(32, 32), parameter_count=1024,
        compression_quality=0.8, hierarchical_indices=indices,
        metadata=ref.ModelMetadata(model_name=name, original_size_bytes=1000,
                                   compressed_size_bytes=500, compression_ratio=0.5,
                                   quantization_timestamp="2024-01-01T00:00:00"))
