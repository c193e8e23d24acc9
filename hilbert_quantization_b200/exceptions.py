"""Exception tree mirroring hilbert_quantization/exceptions.py.  When the reference
package is importable its own classes are re-used, so `except` clauses and
`pytest.raises` written against the reference keep matching."""
try:  # pragma: no cover - depends on the environment
    from hilbert_quantization.exceptions import (  # type: ignore
        HilbertQuantizationError, HilbertMappingError, IndexGenerationError, SearchError,
        QuantizationError, ValidationError, ConfigurationError)
except Exception:  # reference not installed: same names, same hierarchy
    class HilbertQuantizationError(Exception):
        """Base exception for Hilbert quantization system."""

    class HilbertMappingError(HilbertQuantizationError):
        pass

    class IndexGenerationError(HilbertQuantizationError):
        pass

    class SearchError(HilbertQuantizationError):
        pass

    class QuantizationError(HilbertQuantizationError):
        pass

    class ValidationError(HilbertQuantizationError):
        pass

    class ConfigurationError(HilbertQuantizationError):
        pass
