"""GPU (-m gpu): the BINNED join (binned_join.cu: queries routed to shared-memory sized tiles of the index) against
the oracle, bit-exact as the sorted set of (query_id, target_id) pairs -- the same contract as the general path
(find_overlaps_impl, library/include/binary/algorithm/interval_tree.hpp:306-334).

The path is chosen automatically only for batches of >= 2 Mi queries against an index beyond L2 (config D at full
size: tests/test_gpu_fullsize.py); here ``BCU_BINNED=1`` forces it on oracle-sized inputs and ``BCU_BIN_ROWS``
shrinks the tiles so that small indexes still have many bins, halos and queries reaching past their bin.
"""
import os

import numpy as np
import pytest

from binary_b200 import DeviceIndex, synth
from cases import canonical, env, random_case

pytestmark = pytest.mark.gpu


def dev_join(ix, ql, qh, qg, qid_base=0, cap=None, count_only=False):
    """bcu_join_dev / bcu_query_count_dev with device buffers (the binned path serves only these)."""
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
    d_ql, d_qh = t(ql), t(qh)
    d_qg = t(qg) if qg is not None else None
    gptr = d_qg.data_ptr() if d_qg is not None else 0
    n = ql.size
    stream = torch.cuda.current_stream().cuda_stream
    d_off = torch.full((n + 1,), -1, dtype=torch.int64, device=dev)
    ix.count_dev(n, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), gptr, stream)
    torch.cuda.synchronize()
    off_count = d_off.cpu().numpy().view(np.uint64).copy()
    if count_only:
        return off_count
    total = int(off_count[-1])
    cap = total if cap is None else cap
    d_hq = torch.full((cap + 64,), -1, dtype=torch.int32, device=dev)
    d_ht = torch.full((cap + 64,), -1, dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    d_off.fill_(-1)
    ix.join_dev(n, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(), d_ht.data_ptr(),
                d_total.data_ptr(), gptr, qid_base, stream)
    torch.cuda.synchronize()
    off = d_off.cpu().numpy().view(np.uint64)
    assert np.array_equal(off, off_count) and int(d_total.item()) == total
    hq, ht = d_hq.cpu().numpy(), d_ht.cpu().numpy()
    assert (hq[cap:] == -1).all() and (ht[cap:] == -1).all()        # nothing at or beyond the capacity
    return off, hq[:min(cap, total)].view(np.uint32), ht[:min(cap, total)].view(np.uint32)


def check_case(c, port, rows=None, window=None, qid_base=0, must_bin=True):
    f = port.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    kw = {}
    if rows:
        kw["BCU_BIN_ROWS"] = rows
    if window:
        kw["BCU_BINNED_COVER"] = window
    with env(BCU_BINNED=1, **kw):
        ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
        info = ix.info()
        assert info["binned_tiles"] > 0 or not must_bin, info   # not eligible -> the general path answers
        off, hq, ht = dev_join(ix, c["ql"], c["qh"], c["qg"], qid_base=qid_base)
    assert np.array_equal(off, want_off)
    counts = np.diff(want_off).astype(np.int64)
    assert np.array_equal(hq, np.repeat(np.arange(counts.size, dtype=np.uint32), counts) + np.uint32(qid_base))
    assert np.array_equal(canonical(off, ht)[1], want_tid)
    ix.close()
    return info, int(want_off[-1])


@pytest.mark.parametrize("seed,kw,rows", [
    (1, dict(n_t=5000, n_q=3000), 256),
    (2, dict(n_t=5000, n_q=9000, n_groups=5, q_groups=7), 128),                      # unknown query groups
    (3, dict(n_t=4000, n_q=2500, inverted_frac=0.3, dup_frac=0.2, extremes=True), 0),   # u32-wide targets: no layout
    (10, dict(n_t=4000, n_q=2500, inverted_frac=0.3, dup_frac=0.2), 512),
    (4, dict(n_t=4000, n_q=5000, long_frac=0.02, n_groups=3), 512),                  # long targets: long coverage lists
    (5, dict(n_t=1, n_q=1), None),
    (6, dict(n_t=3, n_q=4097, span=50, max_len=10), None),                            # tile boundary + 1
    (7, dict(n_t=300, n_q=4096, span=100000, max_len=100), 64),
    (8, dict(n_t=60000, n_q=40000, span=4_000_000_000, max_len=100000, n_groups=200, q_groups=210), 1024),
    (9, dict(n_t=20000, n_q=12000, span=2_000_000, max_len=900), None),               # default tile size
])
def test_binned_path_random_cases(port_oracle, seed, kw, rows):
    check_case(random_case(seed, **kw), port_oracle, rows=rows or None, window=1e9, must_bin=rows != 0)


def test_windows_beyond_the_hit_mask_spill_to_the_general_index(port_oracle):
    """Dense data: most coverage lists exceed the 32-row mask, so nearly every query goes through
    bin_spill_kernel; mixed with sparse queries that stay in the tiles."""
    c = random_case(21, n_t=20000, n_q=6000, span=200000, max_len=3000)               # ~150 candidates per query
    s = random_case(22, n_t=20000, n_q=6000, span=200000, max_len=3)
    c["ql"] = np.concatenate([c["ql"], s["ql"] + 250000])                             # beyond the dense region: few hits
    c["qh"] = np.concatenate([c["qh"], s["qh"] + 250000])
    c["qg"] = None
    info, hits = check_case(c, port_oracle, rows=2048, window=1e9)
    assert hits > 100 * 6000


@pytest.mark.parametrize("name,n_t,n_q,rows", [("B", 200_000, 300_000, 4096), ("D", 400_000, 500_000, None)])
def test_baseline_configs_scaled_binned(port_oracle, name, n_t, n_q, rows):
    w = synth.CONFIGS[name].scaled(n_t, n_q)
    tg, tl, th = w.targets()
    qg, ql, qh = w.queries()
    info, hits = check_case(dict(tl=tl, th=th, tg=tg, ql=ql, qh=qh, qg=qg), port_oracle, rows=rows, qid_base=12345)
    assert hits > 0 and info["binned_tiles"] >= 25


def test_count_mode_and_capacity_on_the_binned_path(port_oracle):
    c = random_case(31, n_t=30000, n_q=20000, n_groups=4, span=3_000_000, max_len=2000, dup_frac=0.05)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    total = int(want_off[-1])
    with env(BCU_BINNED=1, BCU_BIN_ROWS=1024):
        ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
        assert ix.info()["binned_tiles"] > 4
        assert np.array_equal(dev_join(ix, c["ql"], c["qh"], c["qg"], count_only=True), want_off)
        for cap in (0, 1, total // 2, total - 1):
            # offsets and the total stay complete, nothing is written at or beyond the capacity (checked in dev_join)
            off, hq, ht = dev_join(ix, c["ql"], c["qh"], c["qg"], cap=cap)
            assert np.array_equal(off, want_off) and hq.size == cap
        # and with room to spare the pairs are right again
        off, hq, ht = dev_join(ix, c["ql"], c["qh"], c["qg"], cap=total + 1000)
        assert np.array_equal(canonical(off, ht)[1], want_tid)
    ix.close()


def test_binned_equals_general_path_on_the_same_index(port_oracle):
    """Same index object, path switched per call: identical offsets and pair sets; and an imported index image
    rebuilds its bin layout."""
    import torch
    c = random_case(41, n_t=50000, n_q=30000, n_groups=6, span=5_000_000, max_len=3000, long_frac=0.001)
    with env(BCU_BINNED=1, BCU_BIN_ROWS=2048):
        ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    assert ix.info()["binned_tiles"] > 0
    with env(BCU_BINNED=0):
        off0, hq0, ht0 = dev_join(ix, c["ql"], c["qh"], c["qg"])
    with env(BCU_BINNED=1):
        off1, hq1, ht1 = dev_join(ix, c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off0, off1) and np.array_equal(hq0, hq1)
    assert np.array_equal(canonical(off0, ht0)[1], canonical(off1, ht1)[1])
    ix.close()
    # an imported image of a one-class index rebuilds the same layout (with several classes the image does not hold
    # the (group, low) order the layout is built on: such an index answers through the general path)
    c = random_case(42, n_t=50000, n_q=30000, n_groups=6, span=5_000_000, max_len=3000)
    with env(BCU_BINNED=1, BCU_BIN_ROWS=2048):
        ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    assert ix.info()["n_components"] == 1 and ix.info()["binned_tiles"] > 0
    with env(BCU_BINNED=1):
        off1, hq1, ht1 = dev_join(ix, c["ql"], c["qh"], c["qg"])
    dev = torch.device("cuda:0")
    nbytes = ix.image_size()
    image = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ix.export_dev(image.data_ptr(), nbytes, stream)
    with env(BCU_BINNED=1, BCU_BIN_ROWS=2048):
        ix2 = DeviceIndex.import_dev(0, image.data_ptr(), nbytes, stream)
    assert ix2.info() == ix.info()
    with env(BCU_BINNED=1):
        off2, hq2, ht2 = dev_join(ix2, c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off2, off1) and np.array_equal(canonical(off2, ht2)[1], canonical(off1, ht1)[1])
    ix.close(); ix2.close()
