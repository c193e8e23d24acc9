"""Host-side dimension logic (scalar integer arithmetic, no data path).

Mirrors core/dimension_calculator.py:17-128 (PowerOf4DimensionCalculator) and
rag/embedding_generation/generator.py:293-312 (no efficiency gate)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

VALID_DIMENSIONS = [4, 16, 64, 256, 1024, 4096, 16384]      # config.py:18
MIN_EFFICIENCY_RATIO = 0.5                                   # config.py:27
DEFAULT_PADDING_VALUE = 0.0                                  # config.py:21


@dataclass
class PaddingConfig:
    """models.py:23-37"""
    target_dimensions: Tuple[int, int]
    padding_value: float
    padding_positions: List[Tuple[int, int]]
    efficiency_ratio: float


class PowerOf4DimensionCalculator:
    _padding_config_cls = None      # dropin.install() points this at the reference's models.PaddingConfig

    def __init__(self, min_efficiency_ratio: float = MIN_EFFICIENCY_RATIO):
        self.min_efficiency_ratio = min_efficiency_ratio

    def calculate_optimal_dimensions(self, param_count: int) -> Tuple[int, int]:
        if param_count <= 0:
            raise ValueError("Parameter count must be positive")
        side = int(math.sqrt(self._find_nearest_power_of_4(param_count)))
        return (side, side)

    def calculate_padding_strategy(self, param_count: int, target_dims: Tuple[int, int]) -> PaddingConfig:
        width, height = target_dims
        total = width * height
        if total < param_count:
            raise ValueError(f"Target dimensions {target_dims} cannot accommodate {param_count} parameters")
        eff = param_count / total
        if eff < self.min_efficiency_ratio:
            raise ValueError(f"Efficiency ratio {eff:.3f} is below minimum {self.min_efficiency_ratio}")
        return (self._padding_config_cls or PaddingConfig)(target_dims, DEFAULT_PADDING_VALUE, self._padding_positions(param_count, target_dims), eff)

    @staticmethod
    def _find_nearest_power_of_4(value: int) -> int:
        if value <= 0:
            return 4
        for size in VALID_DIMENSIONS:
            if size >= value:
                return size
        p = VALID_DIMENSIONS[-1]
        while p < value:
            p *= 4
        return p

    @staticmethod
    def _padding_positions(param_count: int, dims: Tuple[int, int]) -> List[Tuple[int, int]]:
        """(x, y) of the unused tail, last cell first (core/dimension_calculator.py:130-156)."""
        width, height = dims
        total = width * height
        return [((total - 1 - i) % width, (total - 1 - i) // width) for i in range(total - param_count)]


def rag_optimal_dimensions(embedding_size: int) -> Tuple[int, int]:
    side = math.ceil(math.sqrt(embedding_size))
    p = 1
    while p < side:
        p *= 2
    while p * p < embedding_size:
        p *= 2
    return (p, p)
