"""Quick kernel timing (dev loop): fused join / count / scatter on a synthetic config, CUDA events, L2 flushed.
usage: python tools/kbench.py [B|C|D] [n_queries] [steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from binary_b200 import DeviceIndex, synth

name = sys.argv[1] if len(sys.argv) > 1 else "B"
w = synth.CONFIGS[name]
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else min(w.n_queries, 10_000_000)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda:0")
tg, tl, th = w.targets(); qg, ql, qh = w.queries(0, n_q)
t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
d_tg, d_tl, d_th, d_qg, d_ql, d_qh = map(t, (tg, tl, th, qg, ql, qh))
stream = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize(); t0 = time.perf_counter()
ix = DeviceIndex.build_dev(tl.size, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr(), stream=stream)
torch.cuda.synchronize(); t1 = time.perf_counter()
ix2 = DeviceIndex.build_dev(tl.size, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr(), stream=stream)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"build: first {1e3*(t1-t0):.1f} ms, second {1e3*(t2-t1):.1f} ms, info {ix.info()}")
d_off = torch.empty(n_q + 1, dtype=torch.int64, device=dev)
ix.count_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), d_qg.data_ptr(), stream)
torch.cuda.synchronize()
hits = int(d_off[-1].item()); cap = hits + 16
d_hq = torch.empty(cap, dtype=torch.int32, device=dev); d_ht = torch.empty(cap, dtype=torch.int32, device=dev)
d_total = torch.zeros(1, dtype=torch.int64, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
alg = synth.algorithmic_bytes(n_q, tl.size, hits)
def timeit(fn, label):
    for _ in range(3): flush.zero_(); fn()
    ms = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    m = float(np.median(ms))
    print(f"{label:8s} median {m*1e3:9.1f} us  min {min(ms)*1e3:9.1f} us  {n_q/m/1e6:8.2f} Gq/s  alg {alg/m/1e6:7.1f} GB/s  frac {alg/m/1e6/6547.5:.3f}")
fused = lambda: ix.join_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(), d_ht.data_ptr(), d_total.data_ptr(), d_qg.data_ptr(), 0, stream)
count = lambda: ix.count_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), d_qg.data_ptr(), stream)
scat = lambda: ix.scatter_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), d_hq.data_ptr(), d_ht.data_ptr(), d_qg.data_ptr(), stream)
print(f"{w.name}: n_q {n_q} hits {hits} ({hits/n_q:.2f}/q) alg bytes {alg/1e6:.1f} MB")
timeit(fused, "fused"); timeit(count, "count"); timeit(scat, "scatter")
assert int(d_total.item()) == hits
