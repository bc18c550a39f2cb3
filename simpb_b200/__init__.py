"""simpb_b200 — SimPB's deformable feature aggregation hot path, B200-native.

Importing this package loads libdfa_b200.so (hand-written sm_100a kernels behind a C ABI).
There is no CPU or eager fallback: a missing library is an ImportError.
"""
from . import cabi  # noqa: F401  (fails loudly when the native library is absent)
from .ops import (  # noqa: F401
    DeformableAggregationFunction,
    deformable_aggregation_function,
    feature_maps_format,
)

__all__ = ["cabi", "DeformableAggregationFunction", "deformable_aggregation_function",
           "feature_maps_format"]
