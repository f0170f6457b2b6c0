"""GPU (-m gpu): memory safety without compute-sanitizer (closed on this pool). ``make -C binary_b200/csrc check``
builds ``libbinary_cuda_check.so`` -- the same sources with ``-DBCU_BOUNDS_CHECK``, whose kernels print and trap on
any index outside the extent of the array it addresses (common.cuh BCU_DEV_ASSERT: directory entries, rows incl.
the 4-row padding the 128-bit loads rely on, probe state, pair and staging capacity, tile offsets in shared memory).
The parity shapes that stress array ends run through it in a subprocess (a trap poisons the CUDA context)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(ROOT, "binary_b200", "libbinary_cuda_check.so")

_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
import oracle
from binary_b200 import DeviceIndex, _lib
from cases import canonical, random_case
from test_gpu_binned import dev_join
assert _lib.LIB_PATH.endswith("libbinary_cuda_check.so")
port = oracle.Oracle("port")
shapes = [
    dict(n_t=5000, n_q=3000), dict(n_t=1, n_q=1), dict(n_t=3, n_q=1025, span=50, max_len=10),
    dict(n_t=2049, n_q=7, span=100, max_len=100), dict(n_t=4000, n_q=2500, inverted_frac=0.3, dup_frac=0.2, extremes=True),
    dict(n_t=4000, n_q=2500, long_frac=0.02, n_groups=3), dict(n_t=20000, n_q=5000, span=200000, max_len=60000),
    dict(n_t=60000, n_q=40000, span=4_000_000_000, max_len=100000, n_groups=300, q_groups=310),
    dict(n_t=1023, n_q=4095, span=5000, max_len=40), dict(n_t=1025, n_q=4097, span=5000, max_len=40),
]
for k, kw in enumerate(shapes):
    c = random_case(100 + k, **kw)
    f = port.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    for binned, rows in (("0", None), ("1", "256"), ("1", None)):
        os.environ["BCU_BINNED"] = binned
        os.environ["BCU_BINNED_COVER"] = "1e9"
        if rows: os.environ["BCU_BIN_ROWS"] = rows
        else: os.environ.pop("BCU_BIN_ROWS", None)
        ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
        off, hq, ht = dev_join(ix, c["ql"], c["qh"], c["qg"])
        assert np.array_equal(off, want_off) and np.array_equal(canonical(off, ht)[1], want_tid), (k, binned, rows)
        if binned == "0":
            off2, hq2, ht2 = ix.join(c["ql"], c["qh"], c["qg"])       # host pipeline
            assert np.array_equal(off2, want_off) and np.array_equal(canonical(off2, ht2)[1], want_tid)
            assert np.array_equal(ix.any(c["ql"], c["qh"], c["qg"]), np.diff(want_off) > 0)
            g2 = ix.join_filtered(c["ql"], c["qh"], c["qg"], kind=2, diff=500, use_strand=False)
            # the rules kernels (sv2nl_rules.cu): TRA columns indexed by record / target id, keys, three probes
            n3 = c["ql"].size // 3
            tra = dict(rec_p1=c["ql"][:n3], rec_p2=c["qh"][:n3], tgt_p1=c["tl"], tgt_p2=c["th"], tgt_pos=c["tl"], tgt_end=c["th"])
            key = np.stack([c["ql"][:n3] % 7] * 4, axis=1)
            qg3 = None if c["qg"] is None else c["qg"][:3 * n3]
            if n3:
                ix.sv2nl_join(c["ql"][:3 * n3], c["qh"][:3 * n3], qg3, diff=1000, probes_per_record=3, tra=tra, rec_key=key)
        ix.close()
print("BOUNDS_OK", len(shapes))
"""


def test_parity_shapes_through_the_bounds_checked_build():
    if not os.path.exists(CHECK_LIB):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "binary_b200", "csrc"), "-j8", "check"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    env = dict(os.environ, BINARY_B200_LIB=CHECK_LIB)
    script = _SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=900)
    assert "BCU_BOUNDS_CHECK failed" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "BOUNDS_OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
