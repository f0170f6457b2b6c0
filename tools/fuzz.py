"""Randomised parity fuzz on the GPU: many small cases of random shape through every entry point of the C ABI
against the oracle tree walk. Every third case forces the stab lists of the long-range emit (BCU_LONG_LISTS=1), every
third the binned path with a random tile size (device-buffer join). usage: python tools/fuzz.py [n_cases] [first_seed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from binary_b200 import DeviceIndex
from cases import canonical, random_case
from test_gpu_binned import dev_join

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
port = oracle.Oracle("port")
t0 = time.time(); pairs = 0
for s in range(seed0, seed0 + n_cases):
    r = np.random.default_rng(s)
    span = int(r.choice([50, 3000, 200_000, 5_000_000, 4_000_000_000]))
    kw = dict(n_t=int(r.integers(1, 40_000)), n_q=int(r.integers(1, 12_000)), span=span,
              max_len=int(max(1, span * r.choice([0.0005, 0.01, 0.3]))), n_groups=int(r.choice([1, 1, 3, 25, 400])),
              inverted_frac=float(r.choice([0, 0, 0.02, 0.5])), dup_frac=float(r.choice([0, 0.1])),
              long_frac=float(r.choice([0, 0, 0.001, 0.05])), extremes=bool(r.integers(0, 2)))
    c = random_case(s, **kw)
    if r.integers(0, 4) == 0:  # some inverted queries
        idx = r.choice(c["ql"].size, max(1, c["ql"].size // 10), replace=False)
        c["ql"][idx], c["qh"][idx] = c["qh"][idx].copy(), c["ql"][idx].copy()
    want_off, want_tid = port.build(c["tl"], c["th"], c["tg"]).query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=8)
    variant = s % 3
    for k in ("BCU_LONG_LISTS", "BCU_BINNED", "BCU_BIN_ROWS", "BCU_BINNED_COVER"):
        os.environ.pop(k, None)
    if variant >= 1:
        os.environ["BCU_LONG_LISTS"] = "1"
    if variant == 2:
        os.environ.update(BCU_BINNED="1", BCU_BIN_ROWS=str(int(r.choice([64, 256, 1024, 8192]))), BCU_BINNED_COVER="1e9")
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    off, hq, ht = ix.join(c["ql"], c["qh"], c["qg"])
    ok = np.array_equal(off, want_off) and np.array_equal(canonical(off, ht)[1], want_tid)
    offd, hqd, htd = dev_join(ix, c["ql"], c["qh"], c["qg"])   # device buffers: the binned path when the index has tiles
    ok = ok and np.array_equal(offd, want_off) and np.array_equal(canonical(offd, htd)[1], want_tid)
    off2 = ix.count(c["ql"], c["qh"], c["qg"])
    hq2, ht2 = ix.scatter(c["ql"], c["qh"], off2, c["qg"])
    ok = ok and np.array_equal(off2, want_off) and np.array_equal(canonical(off2, ht2)[1], want_tid) and np.array_equal(hq, hq2)
    ok = ok and np.array_equal(ix.any(c["ql"], c["qh"], c["qg"]), np.diff(want_off) > 0)
    info = ix.info(); ix.close()
    pairs += int(want_off[-1])
    if not ok:
        print(f"MISMATCH seed {s} {kw} index {info}", flush=True); sys.exit(1)
    if (s - seed0 + 1) % 10 == 0:
        print(f"... {s - seed0 + 1} cases OK, {pairs} pairs, {time.time()-t0:.0f} s", flush=True)
print(f"fuzz OK: {n_cases} cases from seed {seed0}, {pairs} pairs compared, {time.time()-t0:.1f} s")
