"""e2e timing of bcu_join with pinned host buffers (dev loop). usage: python tools/e2ebench.py [B|C|D] [n_q]"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from binary_b200 import DeviceIndex, synth, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "B"
w = synth.CONFIGS[name]; n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
tg, tl, th = w.targets(); qg, ql, qh = w.queries(0, n_q)
ix = DeviceIndex.build(tl, th, tg)
lib = _lib.load()
off = ix.count(ql, qh, qg); total = int(off[-1]); cap = total + 1024
def pinned(a, dt):
    t = torch.empty(a if isinstance(a, int) else a.size, dtype=dt).pin_memory()
    if not isinstance(a, int): t.numpy()[:] = a.view(np.int32)
    return t
h_qg, h_ql, h_qh = pinned(qg, torch.int32), pinned(ql, torch.int32), pinned(qh, torch.int32)
h_off, h_hq, h_ht = pinned(n_q + 1, torch.int64), pinned(cap, torch.int32), pinned(cap, torch.int32)
tot = C.c_uint64()
def run():
    _lib.check(lib.bcu_join(ix._h, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(), h_off.data_ptr(), cap,
                            h_hq.data_ptr(), h_ht.data_ptr(), C.byref(tot)))
for _ in range(3): run()
ts = []
for _ in range(10):
    t0 = time.perf_counter(); run(); ts.append(time.perf_counter() - t0)
assert tot.value == total
print(f"{w.name} chunk={os.environ.get('BCU_HOST_CHUNK','default')}: e2e median {np.median(ts)*1e3:.3f} ms min {min(ts)*1e3:.3f} ms -> {n_q/np.median(ts)/1e9:.2f} Gq/s; bytes h2d {12*n_q/1e6:.0f} MB d2h {(8*(n_q+1)+8*total)/1e6:.0f} MB")
