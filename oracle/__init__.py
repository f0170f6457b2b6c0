"""TEST INFRASTRUCTURE ONLY -- the parity oracle for the interval-overlap join.

Nothing under ``binary_b200/`` imports this package. Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may, and only as the checker / the CPU baseline -- never as the product path.

Two interchangeable CPU back ends behind one ctypes wrapper (:class:`Oracle`):

``kind="port"``       ``oracle/liboracle.so``  -- plain-C restatement of the reference's red-black
                      augmented interval tree (``oracle/interval_oracle.c``; cites
                      ``library/include/binary/algorithm/interval_tree.hpp`` and ``rb_tree.hpp``).
``kind="reference"``  ``oracle/_ref/libbinary_ref.so`` -- the UNMODIFIED reference headers compiled
                      from ``/root/reference`` behind a C ABI (``oracle/ref_harness.cpp``). Built in
                      the authoring container only; the prebuilt file travels to the GPU box.

Plus :func:`brute_pairs` (numpy, the bare predicate ``q.low <= t.high && t.low <= q.high``,
interval_tree.hpp:119-121) for small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libbinary_ref.so")

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(verbose: bool = False) -> None:
    """Compile the oracle (``make -C oracle``): the C port always, ``_ref`` when /root/reference exists."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def _as_u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_u32p)


class Oracle:
    """ctypes front end over ``liboracle.so`` (``orc_*``) or ``_ref/libbinary_ref.so`` (``ref_*``)."""

    def __init__(self, kind: str = "port"):
        if kind == "port":
            if not os.path.exists(PORT_SO):
                build()
            self._lib, self._p = C.CDLL(PORT_SO), "orc_"
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(f"{REF_SO} missing (built only where /root/reference exists)")
            self._lib, self._p = C.CDLL(REF_SO), "ref_"
        else:
            raise ValueError(kind)
        self.kind = kind
        f = self._fn
        f("build").restype = C.c_void_p
        f("build").argtypes = [C.c_uint64, _u32p, _u32p, _u32p]
        f("free").restype = None
        f("free").argtypes = [C.c_void_p]
        f("size").restype = C.c_uint64
        f("size").argtypes = [C.c_void_p, C.c_uint32]
        f("root").restype = C.c_int
        f("root").argtypes = [C.c_void_p, C.c_uint32] + [_u32p] * 5
        f("check_invariants").restype = C.c_int
        f("check_invariants").argtypes = [C.c_void_p, C.c_uint32]
        f("find_overlap").restype = C.c_int
        f("find_overlap").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32] + [_u32p] * 3
        f("query").restype = C.c_int
        f("query").argtypes = [C.c_void_p, C.c_uint64, _u32p, _u32p, _u32p, C.c_int, _u64p,
                               C.POINTER(_u32p), C.POINTER(C.c_double)]
        f("free_buf").restype = None
        f("free_buf").argtypes = [C.c_void_p]
        f("hardware_threads").restype = C.c_int
        if kind == "port":
            L = self._lib
            L.orc_brute.restype = C.c_int
            L.orc_brute.argtypes = [C.c_uint64, _u32p, _u32p, _u32p, C.c_uint64, _u32p, _u32p, _u32p,
                                    _u64p, C.POINTER(_u32p)]
            L.orc_flat_count_hash.restype = C.c_int
            L.orc_flat_count_hash.argtypes = [C.c_uint64, _u32p, _u32p, _u32p, C.c_uint64, _u32p, _u32p,
                                              _u32p, C.c_uint64, C.c_int, _u64p, _u64p, _u64p]
            L.orc_pair_hash.restype = C.c_uint64
            L.orc_pair_hash.argtypes = [C.c_uint64, _u32p, _u32p]

    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def hardware_threads(self) -> int:
        return int(self._fn("hardware_threads")())

    # ---- forest = one reference tree per group (sv2nl mapper.hpp:147-162) -------------------------
    def build(self, low, high, group=None) -> "Forest":
        low, high = _as_u32(low), _as_u32(high)
        group = None if group is None else _as_u32(group)
        assert low.shape == high.shape and (group is None or group.shape == low.shape)
        h = self._fn("build")(low.size, _ptr(group), _ptr(low), _ptr(high))
        if not h:
            raise MemoryError("oracle build failed")
        return Forest(self, h, low.size)

    # ---- the bare predicate, C loop (port only) -----------------------------------------------------
    def brute(self, tlow, thigh, qlow, qhigh, tgroup=None, qgroup=None) -> Tuple[np.ndarray, np.ndarray]:
        assert self.kind == "port"
        tlow, thigh, qlow, qhigh = map(_as_u32, (tlow, thigh, qlow, qhigh))
        tg = None if tgroup is None else _as_u32(tgroup)
        qg = None if qgroup is None else _as_u32(qgroup)
        offsets = np.zeros(qlow.size + 1, dtype=np.uint64)
        out = _u32p()
        self._lib.orc_brute(tlow.size, _ptr(tg), _ptr(tlow), _ptr(thigh), qlow.size, _ptr(qg), _ptr(qlow),
                            _ptr(qhigh), offsets.ctypes.data_as(_u64p), C.byref(out))
        n = int(offsets[-1])
        tid = np.ctypeslib.as_array(out, shape=(max(n, 1),))[:n].copy()
        self._lib.orc_free_buf(out)
        return offsets, tid

    def flat_count_hash(self, tlow, thigh, qlow, qhigh, tgroup=None, qgroup=None, qid_base=0,
                        threads=0, want_counts=False):
        """(total hits, order-independent pair hash[, per-query counts]) via the CPU flat-index twin."""
        assert self.kind == "port"
        tlow, thigh, qlow, qhigh = map(_as_u32, (tlow, thigh, qlow, qhigh))
        tg = None if tgroup is None else _as_u32(tgroup)
        qg = None if qgroup is None else _as_u32(qgroup)
        counts = np.zeros(qlow.size, dtype=np.uint64) if want_counts else None
        total, h = C.c_uint64(), C.c_uint64()
        self._lib.orc_flat_count_hash(tlow.size, _ptr(tg), _ptr(tlow), _ptr(thigh), qlow.size, _ptr(qg),
                                      _ptr(qlow), _ptr(qhigh), qid_base, threads or self.hardware_threads(),
                                      None if counts is None else counts.ctypes.data_as(_u64p),
                                      C.byref(total), C.byref(h))
        return (total.value, h.value, counts) if want_counts else (total.value, h.value)

    def pair_hash(self, query_id, target_id) -> int:
        assert self.kind == "port"
        q, t = _as_u32(query_id), _as_u32(target_id)
        return int(self._lib.orc_pair_hash(q.size, _ptr(q), _ptr(t)))


class Forest:
    def __init__(self, oracle: Oracle, handle, n):
        self._o, self._h, self.n = oracle, handle, n

    def __del__(self):
        if getattr(self, "_h", None):
            self._o._fn("free")(self._h)
            self._h = None

    def size(self, group: int = 0) -> int:
        return int(self._o._fn("size")(self._h, group))

    def root(self, group: int = 0):
        """dict(key, low, high, id, max) of the group's root node, or None."""
        v = [C.c_uint32() for _ in range(5)]
        ok = self._o._fn("root")(self._h, group, *[C.byref(x) for x in v])
        return dict(zip(("key", "low", "high", "id", "max"), (x.value for x in v))) if ok else None

    def check_invariants(self, group: int = 0) -> int:
        """black height (>=0); -1 = black-height mismatch; -2 = wrong max augmentation."""
        return int(self._o._fn("check_invariants")(self._h, group))

    def find_overlap(self, qlow: int, qhigh: int, group: int = 0):
        v = [C.c_uint32() for _ in range(3)]
        ok = self._o._fn("find_overlap")(self._h, group, qlow, qhigh, *[C.byref(x) for x in v])
        return tuple(x.value for x in v) if ok else None

    def query(self, qlow, qhigh, qgroup=None, threads: int = 1, want_targets: bool = True):
        """Batched find_overlaps. Returns (offsets u64[n_q+1], target ids u32 in NATIVE order, seconds)."""
        qlow, qhigh = _as_u32(qlow), _as_u32(qhigh)
        qg = None if qgroup is None else _as_u32(qgroup)
        offsets = np.zeros(qlow.size + 1, dtype=np.uint64)
        out, secs = _u32p(), C.c_double()
        rc = self._o._fn("query")(self._h, qlow.size, _ptr(qg), _ptr(qlow), _ptr(qhigh), threads,
                                  offsets.ctypes.data_as(_u64p), C.byref(out) if want_targets else None,
                                  C.byref(secs))
        if rc != 0:
            raise MemoryError("oracle query failed")
        tid = None
        if want_targets:
            n = int(offsets[-1])
            tid = np.ctypeslib.as_array(out, shape=(max(n, 1),))[:n].copy()
            self._o._fn("free_buf")(out)
        return offsets, tid, secs.value

    def query_sorted_pairs(self, qlow, qhigh, qgroup=None, threads: int = 1):
        """The parity contract: (offsets, target ids sorted ascending within each query)."""
        offsets, tid, _ = self.query(qlow, qhigh, qgroup, threads)
        return offsets, sort_within_segments(offsets, tid)


def sort_within_segments(offsets: np.ndarray, tid: np.ndarray) -> np.ndarray:
    """Sort ids ascending inside each CSR segment (canonical (query_id, target_id) order)."""
    if tid.size == 0:
        return tid
    counts = np.diff(offsets).astype(np.int64)
    qid = np.repeat(np.arange(counts.size, dtype=np.uint64), counts)
    order = np.lexsort((tid, qid))
    return tid[order]


def brute_pairs(tlow, thigh, qlow, qhigh, tgroup=None, qgroup=None):
    """numpy O(n_t*n_q) predicate; returns (offsets, sorted target ids). Small inputs only."""
    tlow, thigh, qlow, qhigh = map(_as_u32, (tlow, thigh, qlow, qhigh))
    hit = (qlow[:, None] <= thigh[None, :]) & (tlow[None, :] <= qhigh[:, None])
    if tgroup is not None or qgroup is not None:
        tg = np.zeros(tlow.size, np.uint32) if tgroup is None else _as_u32(tgroup)
        qg = np.zeros(qlow.size, np.uint32) if qgroup is None else _as_u32(qgroup)
        hit &= qg[:, None] == tg[None, :]
    counts = hit.sum(axis=1).astype(np.uint64)
    offsets = np.zeros(qlow.size + 1, dtype=np.uint64)
    np.cumsum(counts, out=offsets[1:])
    tid = np.nonzero(hit)[1].astype(np.uint32)
    return offsets, tid


def pair_hash_np(query_id: np.ndarray, target_id: np.ndarray) -> int:
    """numpy twin of orc_pair_hash: sum(mix64(q<<32|t)) mod 2^64."""
    z = (query_id.astype(np.uint64) << np.uint64(32)) | target_id.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        return int(z.sum(dtype=np.uint64))
