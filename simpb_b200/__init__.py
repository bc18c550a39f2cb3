"""simpb_b200 — SimPB's deformable feature aggregation hot path, B200-native.

Importing this package loads libdfa_b200.so (hand-written sm_100a kernels behind a C ABI).
There is no CPU or eager fallback: a missing library is an ImportError.

    ops       the reference's `ops` package surface (deformable_aggregation_function, feature_maps_format)
    blocks    DeformableFeatureAggregation / SparseBox3DKeyPointsGenerator on the fused kernels
    msda      multi-scale deformable attention of the 2-D query branch
    parallel  batch sharding and the flat gradient bucket of the data-parallel recipe
    cabi      ctypes binding of every C-ABI entry point
"""
from . import cabi  # noqa: F401  (fails loudly when the native library is absent)
from .ops import (  # noqa: F401
    DeformableAggregationFunction,
    deformable_aggregation_function,
    feature_maps_format,
)
from . import blocks, msda, parallel  # noqa: F401,E402

__all__ = ["cabi", "blocks", "msda", "parallel", "DeformableAggregationFunction",
           "deformable_aggregation_function", "feature_maps_format"]
