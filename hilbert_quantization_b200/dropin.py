"""Inject the device classes into an importable `hilbert_quantization` (the reference package).

    import hilbert_quantization_b200.dropin as dropin
    dropin.install()          # before the reference objects are constructed
    ...                       # unchanged reference callers now run the hot path on the GPU
    dropin.uninstall()

For every hot-path class of SURVEY 8(b) a *hybrid* class is built: the device class of this
package first in the MRO, the reference class second.  A hybrid therefore IS-A reference class
(and through it a subclass of the reference's ABCs, interfaces.py:43-226), answers every
method this package implements from the device, and leaves everything that is not on the
path (statistics, validation helpers, caching, document plumbing) to the reference's own
code.  The hybrids replace the reference classes in every loaded `hilbert_quantization.*`
module (identity match, so aliased imports such as core/pipeline.py:19
`HilbertCurveMapper as HilbertMapperImpl` and function-level imports such as
core/streaming_index_builder.py:284 are covered), which makes default-constructed
`QuantizationPipeline()`, `ReconstructionPipeline()`, `HilbertQuantizer()`,
`StreamingHilbertIndexGenerator()` ... use them without touching the callers.

The hybrid of `RAGSearchEngineImpl` runs in `strict` mode: index rows are located with the
reference's own height heuristic (rag/search/engine.py:134-162) instead of the explicit
`original_height` the batched device path carries.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, Tuple

_STATE: Dict[str, object] = {"installed": False, "originals": [], "hybrids": {}}

# (reference module, reference class, attribute of this package, names that stay the reference's)
_TABLE = [
    ("core.hilbert_mapper", "HilbertCurveMapper", "HilbertCurveMapper", ()),
    ("rag.embedding_generation.hilbert_mapper", "HilbertCurveMapperImpl", "HilbertCurveMapperImpl", ()),
    ("core.dimension_calculator", "PowerOf4DimensionCalculator", "PowerOf4DimensionCalculator", ()),
    ("core.streaming_index_builder", "StreamingHilbertIndexGenerator", "StreamingHilbertIndexGenerator", ()),
    ("core.index_generator", "HierarchicalIndexGeneratorImpl", "HierarchicalIndexGeneratorImpl", ()),
    ("rag.embedding_generation.hierarchical_index_generator", "HierarchicalIndexGenerator", "HierarchicalIndexGenerator",
     ("_detect_original_image_height",)),
    ("core.search_engine", "ProgressiveSimilaritySearchEngine", "ProgressiveSimilaritySearchEngine",
     ("_parse_index_structure",)),
    ("core.precomputed_hilbert_index", "PrecomputedHilbertIndexer", "PrecomputedHilbertIndexer", ()),
    ("rag.search.engine", "RAGSearchEngineImpl", "RAGSearchEngineImpl", ("_get_all_candidate_embeddings",)),
]


def _make_hybrid(ours, ref, keep_ref: Tuple[str, ...], device, extra: dict):
    def __init__(self, *args, **kwargs):
        ref.__init__(self, *args, **kwargs)
        self._device = device
        for k, v in extra.items():
            setattr(self, k, v)
    ns = {"__init__": __init__, "__module__": ref.__module__, "__doc__": ref.__doc__, "_hq_reference_class": ref,
          "_hq_device_class": ours}
    for name in keep_ref:
        ns[name] = ref.__dict__[name]
    return type(ref.__name__, (ours, ref), ns)


def install(package: str = "hilbert_quantization", device=None) -> Dict[str, type]:
    """Build the hybrids and rebind them in every loaded module of `package`.  Returns {class name: hybrid}."""
    if _STATE["installed"]:
        return dict(_STATE["hybrids"])          # type: ignore[arg-type]
    import hilbert_quantization_b200 as hq
    root = importlib.import_module(package)
    models = importlib.import_module(package + ".models")
    hybrids: Dict[str, type] = {}
    pairs = []
    for mod_name, cls_name, our_name, keep in _TABLE:
        try:
            mod = importlib.import_module(f"{package}.{mod_name}")
        except Exception:                         # an optional reference module that does not import here
            continue
        ref = getattr(mod, cls_name)
        extra = {}
        if cls_name == "ProgressiveSimilaritySearchEngine":
            extra["_result_cls"] = models.SearchResult
        if cls_name == "PowerOf4DimensionCalculator":
            extra["_padding_config_cls"] = models.PaddingConfig
        if cls_name == "RAGSearchEngineImpl":
            extra["strict"] = True
        if cls_name == "PrecomputedHilbertIndexer":
            extra["_level_cls"], extra["_index_cls"] = mod.PrecomputedLevel, mod.PrecomputedIndex
        hyb = _make_hybrid(getattr(hq, our_name), ref, keep, device, extra)
        hybrids[cls_name] = hyb
        pairs.append((ref, hyb))
    originals = []
    # exceptions raised by the device classes must be the reference's own classes (`pytest.raises`, `except` clauses)
    try:
        ref_exc = importlib.import_module(package + ".exceptions")
        from . import exceptions as our_exc, mapper as our_mapper
        for name in ("HilbertQuantizationError", "HilbertMappingError", "IndexGenerationError", "SearchError",
                     "QuantizationError", "ValidationError", "ConfigurationError"):
            if hasattr(ref_exc, name):
                for mod in (our_exc, our_mapper, hq):
                    if hasattr(mod, name) and getattr(mod, name) is not getattr(ref_exc, name):
                        originals.append((mod, name, getattr(mod, name)))
                        setattr(mod, name, getattr(ref_exc, name))
        for cls in (our_mapper._MapperCore, our_mapper.HilbertCurveMapper):
            if cls.__dict__.get("_exc") is not None and cls._exc is not ref_exc.HilbertQuantizationError \
                    and cls._exc is not ValueError:
                originals.append((cls, "_exc", cls._exc))
                cls._exc = ref_exc.HilbertQuantizationError
    except ImportError:
        pass
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == package or name.startswith(package + ".")):
            continue
        for attr, val in list(vars(mod).items()):
            for ref, hyb in pairs:
                if val is ref:
                    setattr(mod, attr, hyb)
                    originals.append((mod, attr, ref))
    _STATE.update(installed=True, originals=originals, hybrids=hybrids, root=root)
    return dict(hybrids)


def uninstall() -> None:
    for mod, attr, ref in _STATE["originals"]:   # type: ignore[union-attr]
        setattr(mod, attr, ref)
    _STATE.update(installed=False, originals=[], hybrids={})


def hybrids() -> Dict[str, type]:
    return dict(_STATE["hybrids"])               # type: ignore[arg-type]
