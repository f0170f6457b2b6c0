"""GPU (-m gpu): the CUDA path through the C ABI vs the oracle, bit-exact as the sorted set of
(query_id, target_id) pairs. Integer work: no tolerance anywhere."""
import json
import os

import numpy as np
import pytest

import oracle
from binary_b200 import DeviceIndex, IntervalTree, synth
from cases import CLRS_NODES, canonical, clrs_arrays, random_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _check_all_entry_points(c, port):
    """count / scatter / fused join / any, all against the oracle tree walk."""
    f = port.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    assert len(ix) == c["tl"].size
    # two-call ABI
    off = ix.count(c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off, want_off)
    hq, ht = ix.scatter(c["ql"], c["qh"], off, c["qg"])
    counts = np.diff(want_off).astype(np.int64)
    want_q = np.repeat(np.arange(counts.size, dtype=np.uint32), counts)
    assert np.array_equal(hq, want_q)
    assert np.array_equal(canonical(off, ht)[1], want_tid)
    # fused single pass
    off2, hq2, ht2 = ix.join(c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off2, want_off) and np.array_equal(hq2, want_q)
    assert np.array_equal(canonical(off2, ht2)[1], want_tid)
    off3, hq3, ht3 = ix.join(c["ql"], c["qh"], c["qg"], want_query_ids=False)   # hit_query = NULL
    assert hq3 is None and np.array_equal(off3, want_off) and np.array_equal(canonical(off3, ht3)[1], want_tid)
    # any-overlap bit (shape-independent part of find_overlap)
    assert np.array_equal(ix.any(c["ql"], c["qh"], c["qg"]), counts > 0)
    ix.close()
    return int(want_off[-1])


def test_reference_known_answers_through_the_drop_in_api():
    # test_interval_tree.cpp:87-137,146-155 re-expressed against the GPU path
    t = IntervalTree()
    t.insert_node(CLRS_NODES)
    assert t.size() == 10 and not t.empty()
    assert len(t.find_overlaps(7, 25)) == 8
    assert len(t.find_overlaps(15, 25)) == 5
    assert {(a, b) for a, b, _ in t.find_overlaps(15, 25)} == {(16, 21), (15, 23), (19, 20), (17, 19), (25, 30)}
    assert t.find_overlap(22, 25)[:2] == (15, 23)   # the only overlapping interval
    assert t.find_overlap(100, 111) is None
    d = IntervalTree()
    for _ in range(4):
        d.insert_node(1, 4)
    assert d.size() == 4 and len(d.find_overlaps(2, 5)) == 4
    s = IntervalTree()
    lo = np.arange(0, 1000, 2, dtype=np.uint32)
    s.insert_node(lo, lo + 3)
    assert s.size() == 500
    assert len(s.find_overlaps(10, 10)) == 2


def test_golden_fixtures_from_reference():
    for name in sorted(os.listdir(GOLDEN)):
        if not (name.startswith("ref_") and name.endswith(".json")):
            continue
        g = json.load(open(os.path.join(GOLDEN, name)))
        arr = lambda k: None if g[k] is None else np.array(g[k], np.uint32)
        ix = DeviceIndex.build(arr("tl"), arr("th"), arr("tg"))
        off, hq, ht = ix.join(arr("ql"), arr("qh"), arr("qg"))
        want_off = np.array(g["offsets"], np.uint64)
        assert np.array_equal(off, want_off), name
        assert np.array_equal(canonical(off, ht)[1], canonical(want_off, np.array(g["targets_native"], np.uint32))[1]), name


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_t=5000, n_q=3000)),
    (2, dict(n_t=5000, n_q=3000, n_groups=5, q_groups=7)),
    (3, dict(n_t=4000, n_q=2500, inverted_frac=0.3, dup_frac=0.2, extremes=True)),
    (4, dict(n_t=4000, n_q=2500, long_frac=0.02, n_groups=3)),
    (5, dict(n_t=1, n_q=1)),
    (6, dict(n_t=3, n_q=1025, span=50, max_len=10)),
    (7, dict(n_t=2049, n_q=7, span=100, max_len=100)),          # every query hits ~everything
    (8, dict(n_t=60000, n_q=40000, span=4_000_000_000, max_len=100000, n_groups=300, q_groups=310)),
    (9, dict(n_t=20000, n_q=5000, span=200000, max_len=60000)),  # dense: warp-cooperative path
])
def test_random_cases_all_entry_points(port_oracle, seed, kw):
    _check_all_entry_points(random_case(seed, **kw), port_oracle)


def test_empty_inputs(port_oracle):
    e = np.empty(0, np.uint32)
    ix = DeviceIndex.build(e, e)
    assert len(ix) == 0
    off, hq, ht = ix.join(np.array([1, 2], np.uint32), np.array([3, 4], np.uint32))
    assert list(off) == [0, 0, 0] and hq.size == 0 and ht.size == 0
    assert list(ix.count([5], [9])) == [0, 0]
    ix2 = DeviceIndex.build(np.array([1], np.uint32), np.array([2], np.uint32))
    off, hq, ht = ix2.join(e, e)
    assert list(off) == [0] and hq.size == 0
    assert ix2.any(e, e).size == 0


def test_pair_capacity_is_reported():
    from binary_b200 import _lib
    import ctypes as C
    lo = np.zeros(100, np.uint32)
    hi = np.full(100, 10, np.uint32)
    ix = DeviceIndex.build(lo, hi)
    ql = np.zeros(10, np.uint32)
    off = np.empty(11, np.uint64)
    hq = np.empty(5, np.uint32)
    ht = np.empty(5, np.uint32)
    total = C.c_uint64()
    rc = _lib.load().bcu_join(ix._h, 10, None, ql.ctypes.data, ql.ctypes.data, off.ctypes.data, 5,
                              hq.ctypes.data, ht.ctypes.data, C.byref(total))
    assert rc == _lib.BCU_E_CAPACITY and total.value == 1000
    assert list(off) == [100 * i for i in range(11)]
    off2, hq2, ht2 = ix.join(ql, ql, pair_capacity=5)   # wrapper grows and retries
    assert hq2.size == 1000


@pytest.mark.parametrize("name,n_t,n_q", [("B", 200_000, 100_000), ("C", 200_000, 20_000), ("D", 400_000, 100_000)])
def test_baseline_configs_scaled_vs_tree_oracle(port_oracle, name, n_t, n_q):
    """BASELINE.json configs 2-4 at oracle-friendly sizes, same laws/seeds as the bench."""
    w = synth.CONFIGS[name].scaled(n_t, n_q)
    tg, tl, th = w.targets()
    qg, ql, qh = w.queries()
    hits = _check_all_entry_points(dict(tl=tl, th=th, tg=tg, ql=ql, qh=qh, qg=qg), port_oracle)
    assert hits > 0


def test_against_unmodified_reference(ref_oracle):
    """Where the prebuilt reference travels with the repo: GPU vs the reference headers themselves."""
    c = random_case(31, n_t=30000, n_q=20000, n_groups=6, inverted_frac=0.05, dup_frac=0.05, extremes=True,
                    long_frac=0.002)
    f = ref_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    off, hq, ht = ix.join(c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off, want_off)
    assert np.array_equal(canonical(off, ht)[1], want_tid)


@pytest.mark.parametrize("name", ["B", "C"])
def test_full_size_configs_count_and_hash(port_oracle, name):
    """BASELINE.json's full sizes (1M x 10M): total hit count, per-query counts and an order-independent
    64-bit hash of all pairs vs the CPU flat-index twin (itself proven equal to the tree walk)."""
    import torch
    w = synth.CONFIGS[name]
    n_q = w.n_queries if name == "B" else 2_000_000   # C at 10M is 6.6 GB of pairs; 2M keeps CPU time sane
    tg, tl, th = w.targets()
    qg, ql, qh = w.queries(0, n_q)
    want_total, want_hash, want_counts = port_oracle.flat_count_hash(tl, th, ql, qh, tg, qg, want_counts=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
    d_tg, d_tl, d_th, d_qg, d_ql, d_qh = map(t, (tg, tl, th, qg, ql, qh))
    ix = DeviceIndex.build_dev(tl.size, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr())
    d_off = torch.empty(n_q + 1, dtype=torch.int64, device=dev)
    cap = want_total + 16
    d_hq = torch.empty(cap, dtype=torch.int32, device=dev)
    d_ht = torch.empty(cap, dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ix.join_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(), d_ht.data_ptr(),
                d_total.data_ptr(), d_qg.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    assert int(d_total.item()) == want_total
    off = d_off.cpu().numpy().view(np.uint64)
    assert np.array_equal(np.diff(off), want_counts)
    hq = d_hq[:want_total].cpu().numpy().view(np.uint32)
    ht = d_ht[:want_total].cpu().numpy().view(np.uint32)
    assert port_oracle.pair_hash(hq, ht) == want_hash
    # two-call path on the same data: count == fused offsets; scatter == same pair multiset
    d_off2 = torch.empty_like(d_off)
    ix.count_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off2.data_ptr(), d_qg.data_ptr(), stream)
    d_hq.zero_(); d_ht.zero_()
    ix.scatter_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off2.data_ptr(), d_hq.data_ptr(), d_ht.data_ptr(),
                   d_qg.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(d_off, d_off2)
    assert port_oracle.pair_hash(d_hq[:want_total].cpu().numpy().view(np.uint32),
                                 d_ht[:want_total].cpu().numpy().view(np.uint32)) == want_hash
    # size-independent property: queries are independent -> any prefix of the batch gives a prefix
    half = n_q // 2
    d_off3 = torch.empty(half + 1, dtype=torch.int64, device=dev)
    ix.count_dev(half, d_ql.data_ptr(), d_qh.data_ptr(), d_off3.data_ptr(), d_qg.data_ptr(), stream)
    torch.cuda.synchronize()
    assert torch.equal(d_off3, d_off[: half + 1])


def test_sharded_join_on_gpu(port_oracle):
    """The N>1 path on one GPU: 3 emulated ranks run their query ranges one after the other with
    query_id_base (as bench.py's ranks do), pieces are concatenated -- must equal the unsharded oracle."""
    import torch
    from binary_b200.sharding import ShardedJoin, assemble
    c = random_case(55, n_t=30000, n_q=20011, n_groups=5, dup_frac=0.05, long_frac=0.001)
    dev = torch.device("cuda:0")
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    stream = torch.cuda.current_stream().cuda_stream

    def gpu_join(ql, qh, qg, qid_base):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
        d_ql, d_qh, d_qg = t(ql), t(qh), t(qg)
        n = ql.size
        d_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
        ix.count_dev(n, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), d_qg.data_ptr(), stream)
        total = int(d_off[-1].item())
        d_hq = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        d_ht = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        d_total = torch.zeros(1, dtype=torch.int64, device=dev)
        ix.join_dev(n, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), total, d_hq.data_ptr(), d_ht.data_ptr(),
                    d_total.data_ptr(), d_qg.data_ptr(), qid_base, stream)
        torch.cuda.synchronize()
        assert int(d_total.item()) == total
        return (d_off.cpu().numpy().view(np.uint64), d_hq[:total].cpu().numpy().view(np.uint32),
                d_ht[:total].cpu().numpy().view(np.uint32))

    pieces = [ShardedJoin(gpu_join, r, 3).run(c["ql"], c["qh"], c["qg"]) for r in (2, 0, 1)]
    off, hq, ht = assemble(pieces)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    assert np.array_equal(off, want_off)
    assert np.array_equal(hq, np.repeat(np.arange(c["ql"].size, dtype=np.uint32), np.diff(want_off).astype(np.int64)))
    assert np.array_equal(canonical(off, ht)[1], want_tid)


def test_length_classes_keep_giant_intervals_from_poisoning_the_index(port_oracle):
    """The classic AIList problem: a few chromosome-sized targets among many short ones
    would drag the running max-end along and make every later query scan from them. The build splits
    targets into length classes; results must not change and the index must report > 1 class."""
    rng = np.random.default_rng(99)
    n_t, n_q = 200_000, 100_000
    tl = rng.integers(0, 100_000_000, n_t).astype(np.uint32)
    th = (tl + rng.integers(10, 3000, n_t)).astype(np.uint32)
    tl[:6] = [5, 1000, 2_000_000, 40_000_000, 0, 70_000_000]
    th[:6] = [99_000_000, 60_000_000, 99_999_999, 41_000_000, 100_000_000, 70_050_000]
    tg = (rng.integers(0, 3, n_t)).astype(np.uint32)
    ql = rng.integers(0, 100_000_000, n_q).astype(np.uint32)
    qh = (ql + rng.integers(0, 500, n_q)).astype(np.uint32)
    qg = (rng.integers(0, 4, n_q)).astype(np.uint32)
    ix = DeviceIndex.build(tl, th, tg)
    assert ix.info()["n_components"] > 1
    c = dict(tl=tl, th=th, tg=tg, ql=ql, qh=qh, qg=qg)
    _check_all_entry_points(c, port_oracle)
    # and with the decomposition switched off the answer is the same (results never depend on it)
    os.environ["BCU_MAX_CLASSES"] = "1"
    try:
        ix1 = DeviceIndex.build(tl, th, tg)
        assert ix1.info()["n_components"] == 1
        off1, _, ht1 = ix1.join(ql[:2000], qh[:2000], qg[:2000])
    finally:
        del os.environ["BCU_MAX_CLASSES"]
    off, _, ht = ix.join(ql[:2000], qh[:2000], qg[:2000])
    assert np.array_equal(off, off1) and np.array_equal(canonical(off, ht)[1], canonical(off1, ht1)[1])


@pytest.mark.parametrize("label,groups", [
    ("sparse-ids-smem-bsearch", [7, 5000, 70000, 2**31, 2**32 - 1]),       # <=256 groups, values >= 1024
    ("many-groups-global-table", list(range(0, 3000, 3))),                  # 1000 groups: descriptors in global memory
])
def test_group_lookup_modes(port_oracle, label, groups):
    """The three group-lookup paths of join.cu (direct map / shared-memory binary search / global table),
    with and without length classes, plus queries for groups that have no target."""
    rng = np.random.default_rng(len(groups))
    n_t, n_q = 30000, 20000
    gv = np.array(groups, dtype=np.uint64)
    tg = gv[rng.integers(0, gv.size, n_t)].astype(np.uint32)
    tl = rng.integers(0, 500_000, n_t).astype(np.uint32)
    th = (tl + rng.integers(0, 400, n_t)).astype(np.uint32)
    th[:20] = tl[:20] + 300_000                                              # a few giants -> length classes
    qpool = np.concatenate([gv, np.array([1, 4999, 123456789], dtype=np.uint64)])   # unknown groups too
    qg = qpool[rng.integers(0, qpool.size, n_q)].astype(np.uint32)
    ql = rng.integers(0, 500_000, n_q).astype(np.uint32)
    qh = (ql + rng.integers(0, 300, n_q)).astype(np.uint32)
    _check_all_entry_points(dict(tl=tl, th=th, tg=tg, ql=ql, qh=qh, qg=qg), port_oracle)


def test_concurrent_host_threads_share_one_index(port_oracle):
    """The ABI promises that an immutable index may be queried from several host threads at once (the
    reference shares one tree across its pool threads, sv2nl mapper.cpp:136-140). ctypes drops the GIL
    during the calls, so these really overlap; each thread has its own staging context and streams."""
    import threading
    c = random_case(66, n_t=50000, n_q=40000, n_groups=4, long_frac=0.001)
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    results, errors = {}, []

    def worker(k):
        try:
            lo, hi = k * 10000, (k + 1) * 10000
            for rep in range(3):
                off, hq, ht = ix.join(c["ql"][lo:hi], c["qh"][lo:hi], c["qg"][lo:hi])
                cnt = ix.count(c["ql"][lo:hi], c["qh"][lo:hi], c["qg"][lo:hi])
                assert np.array_equal(off, cnt)
            results[k] = (off, ht)
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for k in range(4):
        lo, hi = k * 10000, (k + 1) * 10000
        want_off, want_tid = f.query_sorted_pairs(c["ql"][lo:hi], c["qh"][lo:hi], c["qg"][lo:hi])
        off, ht = results[k]
        assert np.array_equal(off, want_off) and np.array_equal(canonical(off, ht)[1], want_tid)


def _host_filter(kind, diff, use_strand, ql, qh, tl, th, strand):
    """numpy twin of sv2nl's check_condition (mapper.cpp:50-79) over pair columns."""
    ql, qh, tl, th = (x.astype(np.int64) for x in (ql, qh, tl, th))
    t_has_q = (tl <= ql) & (th >= qh)
    near = (np.abs(ql - tl) <= diff) & (np.abs(qh - th) <= diff)
    if kind == 1:
        return t_has_q & near
    q_has_t = (ql <= tl) & (qh >= th)
    ok = ~t_has_q & ~q_has_t & near
    if use_strand:
        s1, s2 = (strand & 1).astype(bool), (strand & 2).astype(bool)
        ok &= np.where(ql <= tl, s1 & ~s2, ~s1 & s2)
    return ok


@pytest.mark.parametrize("kind,use_strand", [(1, True), (2, True), (2, False)])
@pytest.mark.parametrize("seed,kw", [
    (41, dict(n_t=20000, n_q=15000, span=400000, max_len=3000, n_groups=4)),                    # short ranges
    (42, dict(n_t=20000, n_q=6000, span=200000, max_len=60000, n_groups=2)),                     # long ranges
    (43, dict(n_t=20000, n_q=8000, span=3_000_000, max_len=2000, long_frac=0.01, n_groups=3)),   # length classes
])
def test_fused_pair_filters(port_oracle, kind, use_strand, seed, kw):
    """bcu_join_filtered == (overlap join of the oracle) filtered with check_condition on the host."""
    c = random_case(seed, **kw)
    rng = np.random.default_rng(seed)
    strand = rng.integers(0, 4, c["ql"].size).astype(np.uint8)
    diff = 2500
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    off, tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    qid = np.repeat(np.arange(c["ql"].size, dtype=np.uint32), np.diff(off).astype(np.int64))
    keep = _host_filter(kind, diff, use_strand, c["ql"][qid], c["qh"][qid], c["tl"][tid], c["th"][tid], strand[qid])
    want_q, want_t = qid[keep], tid[keep]
    want_off = np.zeros(c["ql"].size + 1, np.uint64)
    np.cumsum(np.bincount(want_q, minlength=c["ql"].size), out=want_off[1:])
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    goff, ghq, ght = ix.join_filtered(c["ql"], c["qh"], c["qg"], kind=kind, diff=diff, use_strand=use_strand,
                                      qstrand=strand)
    assert 0 < want_q.size < qid.size
    assert np.array_equal(goff, want_off) and np.array_equal(ghq, want_q)
    assert np.array_equal(canonical(goff, ght)[1], canonical(want_off, want_t)[1])


@pytest.mark.parametrize("seed,kw", [
    (91, dict(n_t=30000, n_q=20000, n_groups=5, dup_frac=0.05, long_frac=0.002)),   # several length classes
    (92, dict(n_t=5000, n_q=3000, n_groups=300, inverted_frac=0.1)),                 # binary-search groups
    (93, dict(n_t=0, n_q=100, n_groups=1)),                                          # empty index
])
def test_index_image_round_trip(port_oracle, seed, kw):
    """bcu_index_export_dev -> bcu_index_import_dev (the build-once/broadcast path of SURVEY 8f.4): the
    imported index answers exactly like the one it was exported from, also after the source is freed."""
    import torch
    c = random_case(seed, **kw)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    src = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    nbytes = src.image_size()
    assert nbytes >= 256 and nbytes % 256 == 0
    image = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    with pytest.raises(RuntimeError):
        src.export_dev(image.data_ptr(), nbytes - 1, stream)       # BCU_E_CAPACITY
    src.export_dev(image.data_ptr(), nbytes, stream)
    want = src.join(c["ql"], c["qh"], c["qg"])
    info = src.info()
    src.close()
    moved = image.clone()                                            # what a broadcast delivers
    del image
    ix = DeviceIndex.import_dev(0, moved.data_ptr(), nbytes, stream)
    del moved                                                        # the index owns copies
    torch.cuda.synchronize()
    got_info = ix.info()
    if info["n_components"] > 1:   # the bin layout needs the (group, low) order, which such an image does not hold
        assert got_info["binned_tiles"] == 0
        got_info["binned_tiles"] = info["binned_tiles"]
    assert got_info == info and len(ix) == c["tl"].size
    got = ix.join(c["ql"], c["qh"], c["qg"])
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    assert np.array_equal(got[0], want_off) and np.array_equal(canonical(got[0], got[2])[1], want_tid)
    assert np.array_equal(ix.any(c["ql"], c["qh"], c["qg"]), np.diff(want_off) > 0)
    garbage = torch.zeros(4096, dtype=torch.uint8, device=dev)
    with pytest.raises(RuntimeError):
        DeviceIndex.import_dev(0, garbage.data_ptr(), 4096, stream)  # not an image
    ix.close()


@pytest.mark.parametrize("label", ["proper", "inverted_targets_in_some_groups", "inverted_queries", "ties",
                                   "length_classes", "point_targets"])
def test_long_ranges_rank_count_and_its_fallbacks(port_oracle, label):
    """Long candidate ranges are COUNTED by rank arithmetic (#low <= q.high - #high < q.low) when the
    segment has no inverted row and the query is proper, and scanned otherwise. Both ways, and their mix
    inside one batch, must give the reference's pairs."""
    rng = np.random.default_rng(1234)
    kw = dict(n_t=24000, n_q=6000, span=200000, max_len=60000, n_groups=4)   # ~hundreds of candidates / query
    if label == "ties":
        kw.update(span=3000, max_len=900)        # many equal lows and highs: lower_bound ties in the rank view
    if label == "length_classes":
        kw.update(span=3_000_000, max_len=40000, long_frac=0.01)
    c = random_case(77, **kw)
    if label == "inverted_targets_in_some_groups":   # groups 0 and 1 lose the shortcut, 2 and 3 keep it
        sel = np.flatnonzero(c["tg"] < 2)
        idx = rng.choice(sel, sel.size // 20, replace=False)
        c["th"][idx] = c["tl"][idx] // 3
    if label == "inverted_queries":
        idx = rng.choice(c["ql"].size, c["ql"].size // 5, replace=False)
        c["ql"][idx], c["qh"][idx] = c["qh"][idx].copy(), c["ql"][idx].copy()
    if label == "point_targets":
        c["th"] = c["tl"].copy()
        c["qh"] = np.minimum(c["ql"].astype(np.uint64) + 30000, 0xFFFFFFFF).astype(np.uint32)
    hits = _check_all_entry_points(c, port_oracle)
    assert hits > 30 * c["ql"].size or label in ("inverted_queries", "length_classes", "point_targets")


@pytest.mark.parametrize("label", ["proper", "inverted_rows_and_queries", "length_classes", "groups_unknown_to_index"])
def test_stab_lists_of_the_long_range_emit(port_oracle, label):
    """emit_long_kernel with the stab lists (index_build.cu build_long_lists: rows with low < bin start <= high, per
    coordinate bin) against the plain range scan on the same data and against the oracle; inverted queries keep
    the scan, inverted rows are never listed, an imported image rebuilds the lists."""
    import torch
    from cases import env
    from test_gpu_binned import dev_join
    rng = np.random.default_rng(99)
    kw = dict(n_t=30000, n_q=5000, span=400000, max_len=70000, n_groups=3, dup_frac=0.02)
    if label == "length_classes":
        kw.update(span=4_000_000, max_len=30000, long_frac=0.02)
    if label == "groups_unknown_to_index":
        kw.update(q_groups=5)
    c = random_case(55, **kw)
    if label == "inverted_rows_and_queries":
        idx = rng.choice(c["tl"].size, c["tl"].size // 25, replace=False)
        c["th"][idx] = c["tl"][idx] // 2
        idx = rng.choice(c["ql"].size, c["ql"].size // 6, replace=False)
        c["ql"][idx], c["qh"][idx] = c["qh"][idx].copy(), c["ql"][idx].copy()
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    sizes = {}
    for lists in (0, 1):
        with env(BCU_LONG_LISTS=lists, BCU_BINNED=0):
            ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
            sizes[lists] = ix.info()["device_bytes"]
            off, hq, ht = dev_join(ix, c["ql"], c["qh"], c["qg"], qid_base=7)
            assert np.array_equal(off, want_off) and np.array_equal(canonical(off, ht)[1], want_tid)
            off_h, hq_h, ht_h = ix.join(c["ql"], c["qh"], c["qg"])                     # host-buffer pipeline
            assert np.array_equal(off_h, want_off) and np.array_equal(canonical(off_h, ht_h)[1], want_tid)
            if lists:
                dev = torch.device("cuda:0")
                nbytes = ix.image_size()
                image = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
                stream = torch.cuda.current_stream().cuda_stream
                ix.export_dev(image.data_ptr(), nbytes, stream)
                ix2 = DeviceIndex.import_dev(0, image.data_ptr(), nbytes, stream)
                assert ix2.info()["device_bytes"] == sizes[1]
                off2, _, ht2 = dev_join(ix2, c["ql"], c["qh"], c["qg"])
                assert np.array_equal(off2, want_off) and np.array_equal(canonical(off2, ht2)[1], want_tid)
                ix2.close()
            ix.close()
    assert sizes[1] > sizes[0]                       # the lists were built when asked for, and only then
    assert int(want_off[-1]) > 20 * c["ql"].size or label == "length_classes"


@pytest.mark.parametrize("cap", [0, 1, 37, 5000])
def test_pair_capacity_never_writes_past_the_buffer(port_oracle, cap):
    """bcu_join_dev with a pair buffer smaller than the result: offsets and total are complete, positions
    below the capacity hold correct pairs, nothing at or beyond the capacity is touched (sentinel check) --
    on a batch that mixes short ranges with long ones (emit_kernel and emit_long_kernel both clip)."""
    import torch
    c = random_case(9, n_t=20000, n_q=600, span=200000, max_len=60000)          # long ranges
    s = random_case(10, n_t=20000, n_q=600, span=200000, max_len=30)            # short ranges, same span
    ql = np.stack([c["ql"], s["ql"]], axis=1).reshape(-1).copy()                # interleaved
    qh = np.stack([c["qh"], s["qh"]], axis=1).reshape(-1).copy()
    f = port_oracle.build(c["tl"], c["th"], None)
    want_off, want_tid = f.query_sorted_pairs(ql, qh, None, threads=4)
    total = int(want_off[-1])
    assert total > 5000
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
    d_tl, d_th, d_ql, d_qh = t(c["tl"]), t(c["th"]), t(ql), t(qh)
    stream = torch.cuda.current_stream().cuda_stream
    ix = DeviceIndex.build_dev(c["tl"].size, d_tl.data_ptr(), d_th.data_ptr(), stream=stream)
    n = ql.size
    d_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_hq = torch.full((cap + 4096,), -1, dtype=torch.int32, device=dev)
    d_ht = torch.full((cap + 4096,), -1, dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    ix.join_dev(n, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(), d_ht.data_ptr(),
                d_total.data_ptr(), 0, 0, stream)
    torch.cuda.synchronize()
    assert int(d_total.item()) == total
    assert np.array_equal(d_off.cpu().numpy().view(np.uint64), want_off)
    hq, ht = d_hq.cpu().numpy(), d_ht.cpu().numpy()
    assert (hq[cap:] == -1).all() and (ht[cap:] == -1).all()                    # untouched beyond the capacity
    want_q = np.repeat(np.arange(n, dtype=np.int64), np.diff(want_off).astype(np.int64))
    assert np.array_equal(hq[:cap].astype(np.int64), want_q[:cap])
    for q in np.unique(want_q[:cap]):                                            # per query: a subset, no repeats
        a, b = int(want_off[q]), int(want_off[q + 1])
        got = ht[a:min(b, cap)].view(np.uint32)
        assert np.unique(got).size == got.size and np.isin(got, want_tid[a:b]).all()
        if b <= cap:
            assert np.array_equal(np.sort(got), want_tid[a:b])
    ix.close()
