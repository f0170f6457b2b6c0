"""Host-side mirror of the reference's ``binary::algorithm::tree::IntervalTree`` interface over the
C ABI of ``libbinary_cuda.so``.

Reference interface mirrored (``library/include/binary/algorithm/``):

===============================  ===========================================================
``IntervalTree.insert_node``     ``RbTree::insert_node`` range / args overloads, rb_tree.hpp:111-117,145-149
``IntervalTree.find_overlaps``   ``IntervalTree::find_overlaps``, interval_tree.hpp:161-168 (all hits)
``IntervalTree.find_overlap``    ``IntervalTree::find_overlap``, interval_tree.hpp:152-159 (first hit / None)
``IntervalTree.size / empty``    ``RbTree::size / empty``, rb_tree.hpp:126-129
``find_overlaps_batch``          NEW: the batched entry point the sv2nl loop (mapper.hpp:207-218) uses
===============================  ===========================================================

Differences, by design: intervals carry a ``group`` (sv2nl builds one tree per chromosome,
mapper.hpp:147-162; here one index holds all groups); hits of one query come back sorted by
``(low, id)`` instead of the reference's tree-shape preorder (the parity contract is the sorted set
of pairs); the device index is (re)built lazily on the first query after an insert.

:class:`DeviceIndex` is the thin 1:1 wrapper of the C ABI (host numpy arrays or raw device pointers).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, vp


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class DeviceIndex:
    """One immutable flat index on one GPU (``bcu_index``)."""

    def __init__(self, handle: int, device: int):
        self._h, self.device = handle, device

    # -- construction ---------------------------------------------------------------------------------
    @classmethod
    def build(cls, low, high, group=None, device: int = 0) -> "DeviceIndex":
        lib = _lib.load()
        low, high = _u32(low), _u32(high)
        group = None if group is None else _u32(group)
        if low.shape != high.shape or low.ndim != 1 or (group is not None and group.shape != low.shape):
            raise ValueError("low/high/group must be 1-D arrays of equal length")
        out = vp()
        check(lib.bcu_index_build(device, low.size, _p(group), _p(low), _p(high), C.byref(out)))
        return cls(out.value, device)

    @classmethod
    def build_dev(cls, n: int, d_low: int, d_high: int, d_group: int = 0, device: int = 0,
                  stream: int = 0) -> "DeviceIndex":
        """Build from DEVICE pointers (ints), e.g. ``tensor.data_ptr()``."""
        lib = _lib.load()
        out = vp()
        check(lib.bcu_index_build_dev(device, n, d_group or None, d_low, d_high, stream or None, C.byref(out)))
        return cls(out.value, device)

    # -- index image: build once, ship to the other GPUs (bcu_index_image_size / export / import) --------
    def image_size(self) -> int:
        n = C.c_uint64()
        check(_lib.load().bcu_index_image_size(self._h, C.byref(n)))
        return n.value

    def export_dev(self, d_image: int, nbytes: int, stream: int = 0) -> None:
        """Write the index image into ``nbytes`` of device memory at ``d_image`` (on this index's device)."""
        check(_lib.load().bcu_index_export_dev(self._h, d_image, nbytes, stream or None))

    @classmethod
    def import_dev(cls, device: int, d_image: int, nbytes: int, stream: int = 0) -> "DeviceIndex":
        """New index on ``device`` from an image resident there (the image may be freed afterwards)."""
        out = vp()
        check(_lib.load().bcu_index_import_dev(device, d_image, nbytes, stream or None, C.byref(out)))
        return cls(out.value, device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().bcu_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        n = C.c_uint64()
        check(_lib.load().bcu_index_size(self._h, C.byref(n)))
        return n.value

    def info(self) -> dict:
        inf = _lib.IndexInfo()
        check(_lib.load().bcu_index_get_info(self._h, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in inf._fields_ if k != "reserved"}

    # -- host-buffer queries ----------------------------------------------------------------------------
    @staticmethod
    def _queries(qlow, qhigh, qgroup):
        qlow, qhigh = _u32(qlow), _u32(qhigh)
        qgroup = None if qgroup is None else _u32(qgroup)
        if qlow.shape != qhigh.shape or qlow.ndim != 1 or (qgroup is not None and qgroup.shape != qlow.shape):
            raise ValueError("qlow/qhigh/qgroup must be 1-D arrays of equal length")
        return qlow, qhigh, qgroup

    def count(self, qlow, qhigh, qgroup=None) -> np.ndarray:
        """CSR offsets (u64[n_q+1]) -- ``bcu_query_count``."""
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        offsets = np.empty(qlow.size + 1, dtype=np.uint64)
        total = C.c_uint64()
        check(_lib.load().bcu_query_count(self._h, qlow.size, _p(qgroup), _p(qlow), _p(qhigh),
                                          offsets.ctypes.data, C.byref(total)))
        return offsets

    def scatter(self, qlow, qhigh, offsets, qgroup=None) -> Tuple[np.ndarray, np.ndarray]:
        """(hit_query, hit_target) for offsets from :meth:`count` -- ``bcu_query_scatter``."""
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        total = int(offsets[-1])
        hq = np.empty(total, dtype=np.uint32)
        ht = np.empty(total, dtype=np.uint32)
        check(_lib.load().bcu_query_scatter(self._h, qlow.size, _p(qgroup), _p(qlow), _p(qhigh),
                                            offsets.ctypes.data, hq.ctypes.data, ht.ctypes.data))
        return hq, ht

    def join(self, qlow, qhigh, qgroup=None, pair_capacity: Optional[int] = None, want_query_ids: bool = True):
        """(offsets, hit_query, hit_target) for the whole batch -- ``bcu_join``.

        ``pair_capacity`` defaults to a guess and is grown once on ``BCU_E_CAPACITY``. With
        ``want_query_ids=False`` the (redundant with ``offsets``) query-id column is not produced or copied
        and ``hit_query`` is returned as ``None``.
        """
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        lib = _lib.load()
        offsets = np.empty(qlow.size + 1, dtype=np.uint64)
        cap = int(pair_capacity) if pair_capacity is not None else max(4 * qlow.size, 1 << 16)
        total = C.c_uint64()
        for _ in range(2):
            hq = np.empty(cap, dtype=np.uint32) if want_query_ids else None
            ht = np.empty(cap, dtype=np.uint32)
            rc = lib.bcu_join(self._h, qlow.size, _p(qgroup), _p(qlow), _p(qhigh), offsets.ctypes.data,
                              cap, _p(hq), ht.ctypes.data, C.byref(total))
            if rc == _lib.BCU_E_CAPACITY:
                cap = total.value
                continue
            check(rc)
            return offsets, (hq[: total.value] if want_query_ids else None), ht[: total.value]
        raise _lib.BinaryCudaError(_lib.BCU_E_CAPACITY, "pair capacity still too small after growing")

    def join_filtered(self, qlow, qhigh, qgroup=None, kind: int = _lib.FILTER_SV2NL_DUP, diff: int = 1_000_000,
                      use_strand: bool = True, qstrand=None, pair_capacity: Optional[int] = None):
        """``bcu_join_filtered``: the join with sv2nl's DUP / INV ``check_condition`` applied on the device.
        ``qstrand``: u8 per query, bit0 = strand1 is '+', bit1 = strand2 is '+' (INV with ``use_strand``)."""
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        qs = None if qstrand is None else np.ascontiguousarray(qstrand, dtype=np.uint8)
        lib = _lib.load()
        flt = _lib.Filter(kind, diff, int(bool(use_strand)), 0)
        offsets = np.empty(qlow.size + 1, dtype=np.uint64)
        cap = int(pair_capacity) if pair_capacity is not None else max(2 * qlow.size, 1 << 16)
        total = C.c_uint64()
        for _ in range(2):
            hq = np.empty(cap, dtype=np.uint32)
            ht = np.empty(cap, dtype=np.uint32)
            rc = lib.bcu_join_filtered(self._h, C.byref(flt), qlow.size, _p(qgroup), _p(qlow), _p(qhigh), _p(qs),
                                       offsets.ctypes.data, cap, hq.ctypes.data, ht.ctypes.data, C.byref(total))
            if rc == _lib.BCU_E_CAPACITY:
                cap = total.value
                continue
            check(rc)
            return offsets, hq[: total.value], ht[: total.value]
        raise _lib.BinaryCudaError(_lib.BCU_E_CAPACITY, "pair capacity still too small after growing")

    def sv2nl_join(self, qlow, qhigh, qgroup=None, *, kind: int = _lib.FILTER_NONE, diff: int = 1_000_000,
                   use_strand: bool = True, qstrand=None, probes_per_record: int = 1, tra=None, rec_key=None,
                   pair_capacity: Optional[int] = None):
        """``bcu_sv2nl_join``: one sv2nl mapper on the device -- the (filtered) join, then TraMapper's
        ``check_condition`` (``tra`` = dict of the six breakpoint columns rec_p1, rec_p2, tgt_p1, tgt_p2, tgt_pos,
        tgt_end) and the duplicate-key rule (``rec_key``: [n_rec, 4] u32). Returns per-RECORD (offsets, targets)."""
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        if qlow.size % probes_per_record:
            raise ValueError("the number of queries is not a multiple of probes_per_record")
        n_rec = qlow.size // probes_per_record
        qs = None if qstrand is None else np.ascontiguousarray(qstrand, dtype=np.uint8)
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        keep = []  # the structure holds raw pointers: keep the arrays alive
        rules = _lib.Sv2nlRules(probes_per_record, int(tra is not None), diff, int(rec_key is not None),
                                None, None, None, None, None, None, None)
        if tra is not None:
            for name in ("rec_p1", "rec_p2", "tgt_p1", "tgt_p2", "tgt_pos", "tgt_end"):
                arr = u32(tra[name])
                if arr.size != (n_rec if name.startswith("rec") else len(self)):
                    raise ValueError(f"{name}: wrong length")
                keep.append(arr)
                setattr(rules, name, arr.ctypes.data)
        if rec_key is not None:
            key = u32(rec_key).reshape(-1)
            if key.size != 4 * n_rec:
                raise ValueError("rec_key must hold four words per record")
            keep.append(key)
            rules.rec_key = key.ctypes.data
        flt = _lib.Filter(kind, diff, int(bool(use_strand)), 0)
        offsets = np.empty(n_rec + 1, dtype=np.uint64)
        cap = int(pair_capacity) if pair_capacity is not None else max(2 * qlow.size, 1 << 16)
        total = C.c_uint64()
        lib = _lib.load()
        for _ in range(2):
            ht = np.empty(cap, dtype=np.uint32)
            rc = lib.bcu_sv2nl_join(self._h, C.byref(flt) if kind != _lib.FILTER_NONE else None, C.byref(rules), n_rec,
                                    _p(qgroup), _p(qlow), _p(qhigh), _p(qs), offsets.ctypes.data, cap, ht.ctypes.data,
                                    C.byref(total))
            if rc == _lib.BCU_E_CAPACITY:
                cap = total.value
                continue
            check(rc)
            return offsets, ht[: total.value]
        raise _lib.BinaryCudaError(_lib.BCU_E_CAPACITY, "pair capacity still too small after growing")

    def any(self, qlow, qhigh, qgroup=None) -> np.ndarray:
        qlow, qhigh, qgroup = self._queries(qlow, qhigh, qgroup)
        out = np.empty(qlow.size, dtype=np.uint8)
        check(_lib.load().bcu_query_any(self._h, qlow.size, _p(qgroup), _p(qlow), _p(qhigh), out.ctypes.data))
        return out.astype(bool)

    # -- device-pointer queries (ints; asynchronous on `stream`) -----------------------------------------
    def count_dev(self, n_q, d_qlow, d_qhigh, d_offsets, d_qgroup=0, stream=0) -> None:
        check(_lib.load().bcu_query_count_dev(self._h, n_q, d_qgroup or None, d_qlow, d_qhigh, d_offsets,
                                              stream or None))

    def scatter_dev(self, n_q, d_qlow, d_qhigh, d_offsets, d_hit_query, d_hit_target, d_qgroup=0,
                    stream=0) -> None:
        check(_lib.load().bcu_query_scatter_dev(self._h, n_q, d_qgroup or None, d_qlow, d_qhigh, d_offsets,
                                                d_hit_query, d_hit_target, stream or None))

    def join_dev(self, n_q, d_qlow, d_qhigh, d_offsets, pair_capacity, d_hit_query, d_hit_target, d_total,
                 d_qgroup=0, query_id_base=0, stream=0) -> None:
        check(_lib.load().bcu_join_dev(self._h, n_q, d_qgroup or None, d_qlow, d_qhigh, d_offsets,
                                       pair_capacity, d_hit_query or None, d_hit_target or None,
                                       d_total or None, query_id_base, stream or None))

    def any_dev(self, n_q, d_qlow, d_qhigh, d_any, d_qgroup=0, stream=0) -> None:
        check(_lib.load().bcu_query_any_dev(self._h, n_q, d_qgroup or None, d_qlow, d_qhigh, d_any,
                                            stream or None))


def join_multi(indexes, qlow, qhigh, qgroup=None, pair_capacity: Optional[int] = None, want_query_ids: bool = True,
               counts32: bool = False):
    """``bcu_join_multi``: one call, several GPUs. ``indexes`` = one :class:`DeviceIndex` per device, replicas of
    the same target set; the batch is cut into ``len(indexes)`` contiguous query ranges. Returns
    ``(offsets or counts, hit_query, hit_target)`` exactly as :meth:`DeviceIndex.join` does for the whole batch
    (``counts32=True``: u32 hits per query instead of u64 offsets -- half the bytes over PCIe)."""
    qlow, qhigh, qgroup = DeviceIndex._queries(qlow, qhigh, qgroup)
    lib = _lib.load()
    handles = (vp * len(indexes))(*[ix._h for ix in indexes])
    n = qlow.size
    offsets = None if counts32 else np.empty(n + 1, dtype=np.uint64)
    counts = np.empty(n, dtype=np.uint32) if counts32 else None
    cap = int(pair_capacity) if pair_capacity is not None else max(4 * n, 1 << 16)
    total = C.c_uint64()
    for _ in range(2):
        hq = np.empty(cap, dtype=np.uint32) if want_query_ids else None
        ht = np.empty(cap, dtype=np.uint32)
        rc = lib.bcu_join_multi(handles, len(indexes), n, _p(qgroup), _p(qlow), _p(qhigh), _p(offsets), _p(counts), cap,
                                _p(hq), ht.ctypes.data, C.byref(total))
        if rc == _lib.BCU_E_CAPACITY:
            cap = total.value
            continue
        check(rc)
        return (counts if counts32 else offsets), (hq[: total.value] if want_query_ids else None), ht[: total.value]
    raise _lib.BinaryCudaError(_lib.BCU_E_CAPACITY, "pair capacity still too small after growing")


class IntervalTree:
    """Drop-in for the reference ``IntervalTree<IntervalNode<UIntInterval>>`` with a batched query."""

    def __init__(self, device: int = 0, devices: Optional[Sequence[int]] = None):
        """``devices``: several GPUs for ``find_overlaps_batch`` -- the index is replicated on each and every batch is
        cut into one contiguous query range per device (``bcu_join_multi``); single queries use the first."""
        self.devices = list(devices) if devices else None
        self.device = self.devices[0] if self.devices else device
        self._replicas: list = []
        self._low: list = []
        self._high: list = []
        self._group: list = []
        self._index: Optional[DeviceIndex] = None
        self._arrays = None

    # reference: insert_node(Args&&...) / insert_node(R&& range)
    def insert_node(self, low, high=None, group=0) -> None:
        """``insert_node(low, high[, group])`` or ``insert_node(iterable of (low, high[, group]))``."""
        if high is None:
            for item in low:
                self.insert_node(*item)
            return
        lo, hi = _u32(np.atleast_1d(low)), _u32(np.atleast_1d(high))
        if lo.shape != hi.shape:
            raise ValueError("low and high differ in length")
        g = _u32(np.broadcast_to(np.atleast_1d(group), lo.shape))
        self._low.append(lo)
        self._high.append(hi)
        self._group.append(g)
        self._index = None
        self._replicas = []

    def size(self) -> int:
        return int(sum(a.size for a in self._low))

    def empty(self) -> bool:
        return self.size() == 0

    def _ensure(self) -> DeviceIndex:
        if self._index is None:
            cat = lambda xs: np.concatenate(xs) if xs else np.empty(0, np.uint32)
            self._arrays = (cat(self._low), cat(self._high), cat(self._group))
            self._low, self._high, self._group = [self._arrays[0]], [self._arrays[1]], [self._arrays[2]]
            self._index = DeviceIndex.build(*self._arrays, device=self.device)
            self._replicas = [DeviceIndex.build(*self._arrays, device=d) for d in (self.devices or [])[1:]]
        return self._index

    def find_overlaps_batch(self, qlow, qhigh, qgroup=None):
        """CSR ``(offsets u64[n_q+1], target_ids u32[total])``; ids are insertion ordinals."""
        ix = self._ensure()
        if self.devices:
            offsets, _, target = join_multi([ix] + self._replicas, qlow, qhigh, qgroup, want_query_ids=False)
        else:
            offsets, _, target = ix.join(qlow, qhigh, qgroup, want_query_ids=False)
        return offsets, target

    def find_overlaps(self, low: int, high: int, group: int = 0):
        """All overlapping intervals as ``(low, high, id)`` tuples (reference: vector<interval_type>)."""
        _, target = self.find_overlaps_batch([low], [high], [group])
        lo, hi, _ = self._arrays
        return [(int(lo[t]), int(hi[t]), int(t)) for t in target]

    def find_overlap(self, low: int, high: int, group: int = 0):
        """First overlapping interval in index order, or ``None`` (reference: std::optional)."""
        hits = self.find_overlaps(low, high, group)
        return hits[0] if hits else None
