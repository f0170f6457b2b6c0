"""CPU: the C-ABI library loads, exports every symbol include/binary_cuda.h declares, and fails LOUDLY
(no CPU fallback) when there is no CUDA device. No compute is attempted without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from binary_b200 import _lib
from binary_b200.interval_tree import DeviceIndex, IntervalTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "binary_cuda.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcu_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in binary_cuda.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in binary_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.bcu_version()
    assert isinstance(lib.bcu_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_touching_cuda():
    lib = _lib.load()
    assert lib.bcu_index_build(0, 4, None, None, None, None) == _lib.BCU_E_INVALID
    out = C.c_void_p()
    assert lib.bcu_index_build(0, 4, None, None, None, C.byref(out)) == _lib.BCU_E_INVALID
    assert b"NULL" in lib.bcu_last_error()
    assert lib.bcu_index_build(0, 1 << 33, None, None, None, C.byref(out)) == _lib.BCU_E_LIMIT
    assert lib.bcu_query_count(None, 0, None, None, None, None, None) == _lib.BCU_E_INVALID
    assert lib.bcu_index_size(None, None) == _lib.BCU_E_INVALID
    assert lib.bcu_index_free(None) == _lib.BCU_OK


def _no_gpu():
    n = C.c_int()
    rc = _lib.load().bcu_device_count(C.byref(n))
    return rc != _lib.BCU_OK or n.value == 0


def test_no_cpu_fallback_without_a_device():
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.BinaryCudaError) as ei:
        DeviceIndex.build(np.array([1, 5], np.uint32), np.array([3, 9], np.uint32))
    assert ei.value.status in (_lib.BCU_E_CUDA, _lib.BCU_E_NOMEM)
    t = IntervalTree()
    t.insert_node(16, 21)
    assert t.size() == 1 and not t.empty()
    with pytest.raises(_lib.BinaryCudaError):
        t.find_overlaps(1, 2)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "binary_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libbinary_ref" not in text, f
