"""ctypes binding of ``libbinary_cuda.so`` (C ABI: ``include/binary_cuda.h``).

The product path has NO CPU fallback: if the shared library is missing this module raises, and every
call fails loudly (``BinaryCudaError``) when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BINARY_B200_LIB") or os.path.join(_HERE, "libbinary_cuda.so")

BCU_OK, BCU_E_INVALID, BCU_E_CUDA, BCU_E_NOMEM, BCU_E_CAPACITY, BCU_E_LIMIT = 0, -1, -2, -3, -4, -5

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p


class BinaryCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libbinary_cuda status {status}: {message}")
        self.status = status


class Filter(C.Structure):
    """``bcu_filter``: kind 1 = sv2nl DUP, 2 = sv2nl INV."""
    _fields_ = [("kind", C.c_uint32), ("diff", C.c_uint32), ("use_strand", C.c_uint32), ("reserved", C.c_uint32)]


FILTER_NONE, FILTER_SV2NL_DUP, FILTER_SV2NL_INV = 0, 1, 2


class Sv2nlRules(C.Structure):
    """``bcu_sv2nl_rules``: what follows the join inside one sv2nl mapper (TRA condition, duplicate-key rule)."""
    _fields_ = [("probes_per_record", C.c_uint32), ("tra", C.c_uint32), ("diff", C.c_uint32), ("dedup", C.c_uint32),
                ("rec_p1", vp), ("rec_p2", vp), ("tgt_p1", vp), ("tgt_p2", vp), ("tgt_pos", vp), ("tgt_end", vp),
                ("rec_key", vp)]


class IndexInfo(C.Structure):
    _fields_ = [("n_targets", C.c_uint64), ("n_groups", C.c_uint32), ("n_components", C.c_uint32),
                ("bin_shift", C.c_uint32), ("sort_passes", C.c_uint32), ("n_bins", C.c_uint64),
                ("device_bytes", C.c_uint64), ("device", C.c_int32), ("binned_tiles", C.c_int32)]


# every symbol include/binary_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "bcu_version": (C.c_char_p, []),
    "bcu_last_error": (C.c_char_p, []),
    "bcu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bcu_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "bcu_host_free": (C.c_int, [vp]),
    "bcu_index_build": (C.c_int, [C.c_int, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "bcu_index_build_dev": (C.c_int, [C.c_int, C.c_uint64, vp, vp, vp, vp, C.POINTER(vp)]),
    "bcu_index_free": (C.c_int, [vp]),
    "bcu_index_size": (C.c_int, [vp, u64p]),
    "bcu_index_get_info": (C.c_int, [vp, C.POINTER(IndexInfo)]),
    "bcu_query_count": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, u64p]),
    "bcu_query_scatter": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp]),
    "bcu_join": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, C.c_uint64, vp, vp, u64p]),
    "bcu_trim": (C.c_int, []),
    "bcu_join_multi": (C.c_int, [C.POINTER(vp), C.c_int, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, vp, u64p]),
    "bcu_join_filtered": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, vp, u64p]),
    "bcu_join_filtered_dev": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp]),
    "bcu_query_any": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp]),
    "bcu_query_count_dev": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp]),
    "bcu_query_scatter_dev": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp, vp, vp]),
    "bcu_join_dev": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, C.c_uint32, vp]),
    "bcu_query_any_dev": (C.c_int, [vp, C.c_uint64, vp, vp, vp, vp, vp]),
    "bcu_index_image_size": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "bcu_index_export_dev": (C.c_int, [vp, vp, C.c_uint64, vp]),
    "bcu_index_import_dev": (C.c_int, [C.c_int, vp, C.c_uint64, vp, C.POINTER(vp)]),
    "bcu_sv2nl_join": (C.c_int, [vp, vp, vp, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, u64p]),
    "bcu_launch_count": (C.c_uint64, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libbinary_cuda.so (built by ``make -C binary_b200/csrc`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C binary_b200/csrc` "
                "(there is no CPU fallback for the overlap join)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != BCU_OK:
        raise BinaryCudaError(status, load().bcu_last_error().decode(errors="replace"))
