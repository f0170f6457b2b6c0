"""CPU: pin the oracle against the reference's own known answers and against outputs of the
UNMODIFIED reference headers (oracle/_ref when built here; committed fixtures everywhere)."""
import json
import os

import numpy as np
import pytest

import oracle
from cases import (CLRS_NODES, CLRS_Q_7_25_NATIVE, CLRS_Q_15_25_NATIVE, clrs_arrays, random_case)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _kinds():
    return ["port"] + (["reference"] if oracle.have_reference() else [])


@pytest.fixture(params=_kinds())
def any_oracle(request):
    return oracle.Oracle(request.param)


def test_clrs_fixture_shape(any_oracle):
    # test_interval_tree.cpp:87-99: size 10, root key 16, equal black heights
    lo, hi = clrs_arrays()
    f = any_oracle.build(lo, hi)
    assert f.size() == 10
    assert f.root()["key"] == 16
    assert f.check_invariants() > 0


def test_500_inserts(any_oracle):
    # test_interval_tree.cpp:101-109: [i, i+3] for i in 0,2,..,998
    lo = np.arange(0, 1000, 2, dtype=np.uint32)
    f = any_oracle.build(lo, lo + 3)
    assert f.size() == 500
    assert f.check_invariants() > 0


def test_find_overlap_single(any_oracle):
    # test_interval_tree.cpp:120-129
    lo, hi = clrs_arrays()
    f = any_oracle.build(lo, hi)
    hit = f.find_overlap(22, 25)
    assert hit is not None and hit[:2] == (15, 23)
    assert f.find_overlap(100, 111) is None


def test_find_overlaps_counts_and_native_order(any_oracle):
    # test_interval_tree.cpp:131-137 (8 and 5 hits) + SURVEY 8(c) probe of the native preorder
    lo, hi = clrs_arrays()
    f = any_oracle.build(lo, hi)
    off, tid, _ = f.query([7, 15], [25, 25])
    assert list(off) == [0, 8, 13]
    got = [CLRS_NODES[t] for t in tid]
    assert got[:8] == CLRS_Q_7_25_NATIVE
    assert got[8:] == CLRS_Q_15_25_NATIVE


def test_duplicates_are_distinct_hits(any_oracle):
    # test_interval_tree.cpp:146-155
    f = any_oracle.build([1, 1, 1, 1], [4, 4, 4, 4])
    assert f.size() == 4
    off, tid, _ = f.query([2], [5])
    assert off[-1] == 4 and sorted(tid) == [0, 1, 2, 3]


def test_empty_forest_and_empty_batch(any_oracle):
    f = any_oracle.build(np.empty(0, np.uint32), np.empty(0, np.uint32))
    off, tid, _ = f.query([1, 2], [3, 4])
    assert list(off) == [0, 0, 0] and tid.size == 0
    f2 = any_oracle.build([1], [2])
    off, tid, _ = f2.query(np.empty(0, np.uint32), np.empty(0, np.uint32))
    assert list(off) == [0]


@pytest.mark.parametrize("seed,kw", [
    (1, dict(n_t=3000, n_q=800)),
    (2, dict(n_t=3000, n_q=800, n_groups=5, q_groups=7)),
    (3, dict(n_t=2000, n_q=600, inverted_frac=0.3, dup_frac=0.2, extremes=True)),
    (4, dict(n_t=2000, n_q=600, long_frac=0.02, n_groups=3)),
])
def test_tree_equals_bare_predicate(any_oracle, port_oracle, seed, kw):
    """The reference walk returns exactly {t : q.low <= t.high && t.low <= q.high}."""
    c = random_case(seed, **kw)
    f = any_oracle.build(c["tl"], c["th"], c["tg"])
    off, tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=2)
    boff, btid = oracle.brute_pairs(c["tl"], c["th"], c["ql"], c["qh"], c["tg"], c["qg"])
    assert np.array_equal(off, boff) and np.array_equal(tid, btid)
    coff, ctid = port_oracle.brute(c["tl"], c["th"], c["ql"], c["qh"], c["tg"], c["qg"])
    assert np.array_equal(off, coff) and np.array_equal(tid, ctid)


@pytest.mark.parametrize("seed", [11, 12])
def test_port_matches_reference_native_order(port_oracle, ref_oracle, seed):
    """Where the reference is built: same tree shape (root, black height) and same preorder hit order."""
    c = random_case(seed, n_t=20000, n_q=5000, n_groups=4, inverted_frac=0.05, dup_frac=0.05, extremes=True)
    fp = port_oracle.build(c["tl"], c["th"], c["tg"])
    fr = ref_oracle.build(c["tl"], c["th"], c["tg"])
    for g in range(4):
        assert fp.root(g) == fr.root(g)
        assert fp.check_invariants(g) == fr.check_invariants(g) > 0
    op, tp, _ = fp.query(c["ql"], c["qh"], c["qg"], threads=3)
    orr, tr, _ = fr.query(c["ql"], c["qh"], c["qg"], threads=3)
    assert np.array_equal(op, orr) and np.array_equal(tp, tr)


def test_flat_twin_equals_tree(port_oracle):
    c = random_case(21, n_t=50000, n_q=20000, n_groups=6, q_groups=8, long_frac=0.001, inverted_frac=0.01)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    off, tid, _ = f.query(c["ql"], c["qh"], c["qg"], threads=2)
    total, h, counts = port_oracle.flat_count_hash(c["tl"], c["th"], c["ql"], c["qh"], c["tg"], c["qg"],
                                                   want_counts=True)
    assert total == off[-1]
    assert np.array_equal(counts, np.diff(off))
    qid = np.repeat(np.arange(c["ql"].size, dtype=np.uint32), np.diff(off).astype(np.int64))
    assert h == port_oracle.pair_hash(qid, tid) == oracle.pair_hash_np(qid, tid)


def test_golden_fixtures_from_reference(port_oracle):
    """tests/golden/ref_*.json were produced by tests/golden/make_golden.py running the UNMODIFIED
    reference (oracle/_ref) in the authoring container; the port must reproduce them everywhere."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".json"))
    assert files, "golden fixtures missing"
    for name in files:
        with open(os.path.join(GOLDEN, name)) as fh:
            g = json.load(fh)
        tg = None if g["tg"] is None else np.array(g["tg"], np.uint32)
        qg = None if g["qg"] is None else np.array(g["qg"], np.uint32)
        f = port_oracle.build(np.array(g["tl"], np.uint32), np.array(g["th"], np.uint32), tg)
        off, tid, _ = f.query(np.array(g["ql"], np.uint32), np.array(g["qh"], np.uint32), qg)
        assert list(map(int, off)) == g["offsets"], name
        assert list(map(int, tid)) == g["targets_native"], name
        for grp, root in g["roots"].items():
            assert f.root(int(grp)) == root, name
            assert f.check_invariants(int(grp)) == g["black_height"][grp], name
