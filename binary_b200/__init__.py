"""binary_b200 -- B200-native (sm_100a) interval-overlap join behind BINARY's IntervalTree API.

Scope: the one data-parallel hot path of ylab-hi/BINARY: sort + flat augmented
index + batched overlap join, as hand-written CUDA in ``libbinary_cuda.so`` (C ABI in
``include/binary_cuda.h``). There is no CPU fallback: importing works without a GPU, calling does not.
"""
from .interval_tree import DeviceIndex, IntervalTree, join_multi  # noqa: F401
from ._lib import BinaryCudaError, LIB_PATH  # noqa: F401

__all__ = ["DeviceIndex", "IntervalTree", "join_multi", "BinaryCudaError", "LIB_PATH"]
