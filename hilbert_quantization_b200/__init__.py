"""hilbert_quantization_b200 -- the sm_100a hot path of hilbert-quantization behind the
reference's own class surface.  Importing the package loads libhq_b200.so (built in-tree by
`python -m hilbert_quantization_b200.build`); there is no CPU fallback."""
from ._lib import HQLibraryError, LIB_PATH                                    # noqa: F401  (fails loudly if the library is missing)
from .exceptions import HilbertQuantizationError                              # noqa: F401
from .dimension import PowerOf4DimensionCalculator, rag_optimal_dimensions    # noqa: F401
from .mapper import HilbertCurveMapper, HilbertCurveMapperImpl               # noqa: F401
from .index import (HierarchicalIndexGenerator, HierarchicalIndexGeneratorImpl,   # noqa: F401
                    StreamingHilbertIndexGenerator, index_from_grids, map_and_index, map_parameter_stream)
from .quantize import FrameQuantizer, dequantize_u8_batch, map_index_quantize, quantize_u8_batch  # noqa: F401
from .search import (EmbeddingDatabase, ProgressiveSimilaritySearchEngine, RAGSearchEngineImpl,   # noqa: F401
                     SearchGraph, SearchResult, comprehensive_scores, search_batch, search_stream)
from .rag import DocumentSearchResult, ProgressiveSearchEngine, RAGSystem    # noqa: F401
from .precomputed import PrecomputedHilbertIndexer, PrecomputedIndex, PrecomputedLevel   # noqa: F401
from .frames import EmbeddingFrame, EmbeddingFrameBatch, QuantizedModelBatch   # noqa: F401
from . import video                                                       # noqa: F401
from .distributed import MergePipeline, ShardedSearch, allgather_merge, shard_bounds        # noqa: F401

__version__ = "0.1.0"
