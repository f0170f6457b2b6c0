"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck), no torch.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from binary_b200 import DeviceIndex
from cases import random_case
import oracle

port = oracle.Oracle("port")
cases = [
    dict(seed=1, n_t=3000, n_q=2500, n_groups=5, q_groups=7),                         # sparse, direct group map
    dict(seed=2, n_t=3000, n_q=1100, span=60000, max_len=30000),                        # long ranges (warp path)
    dict(seed=3, n_t=4000, n_q=2100, long_frac=0.01, n_groups=3, inverted_frac=0.1, dup_frac=0.1, extremes=True),
    dict(seed=4, n_t=2000, n_q=900, span=4_000_000_000, max_len=100000, n_groups=300, q_groups=310),
    dict(seed=5, n_t=1, n_q=1),
]
for kw in cases:
    seed = kw.pop("seed")
    c = random_case(seed, **kw)
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = port.build(c["tl"], c["th"], c["tg"]).query_sorted_pairs(c["ql"], c["qh"], c["qg"])
    off = ix.count(c["ql"], c["qh"], c["qg"])
    hq, ht = ix.scatter(c["ql"], c["qh"], off, c["qg"])
    off2, hq2, ht2 = ix.join(c["ql"], c["qh"], c["qg"])
    anyhit = ix.any(c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off, want_off) and np.array_equal(off2, want_off)
    assert np.array_equal(oracle.sort_within_segments(off, ht), want_tid)
    assert np.array_equal(oracle.sort_within_segments(off2, ht2), want_tid)
    assert np.array_equal(anyhit, np.diff(want_off) > 0)
    print("ok", seed, ix.info()["n_components"], int(off[-1]))
    ix.close()
print("ALL OK")
