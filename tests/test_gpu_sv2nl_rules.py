"""GPU (-m gpu): ``bcu_sv2nl_join`` -- the join plus sv2nl's rules that follow it, on the device (csrc/sv2nl_rules.cu)
-- against the oracle join with the same rules applied in numpy / plain Python:

* TraMapper::check_condition (standalone/sv2nl/source/mapper.cpp:144-156) with the raw-interval overlap of the
  reference's tree of unvalidated BND records (mapper.cpp:103,158-170), on the re-keyed join with one and with three
  probes per record;
* the SV2NL_USE_CACHE duplicate-key rule (include/mapper.hpp:204-229): the first record of a key that keeps a pair
  is written, later ones are not;
* the fused DUP / INV filters combined with the duplicate-key rule.

Bit-exact: per record the sorted kept target ids. (The file-level behaviour of the two front ends built on this entry
is checked against the UNMODIFIED reference in tests/test_sv2nl_reference.py.)
"""
import numpy as np
import pytest

from binary_b200 import DeviceIndex, _lib
from cases import random_case

pytestmark = pytest.mark.gpu


def _per_record_sorted(off, tgt):
    return [np.sort(tgt[int(off[r]):int(off[r + 1])]) for r in range(off.size - 1)]


def _dedup_first_with_hits(keys, kept_lists):
    """keys: [n, 4]; a record is written only if no EARLIER record with the same key was written."""
    seen, out = set(), []
    for k, lst in zip(map(tuple, keys.tolist()), kept_lists):
        if lst.size and k not in seen:
            seen.add(k)
            out.append(lst)
        else:
            out.append(lst[:0])
    return out


def _tra_case(seed, n_t, n_rec, n_pairs, span, diff):
    rng = np.random.default_rng(seed)
    t = dict(g=rng.integers(0, n_pairs, n_t), p1=rng.integers(0, span, n_t), p2=rng.integers(0, span, n_t))
    swapped = rng.random(n_t) < 0.5                       # chrom > chr2 in the file: POS is the SECOND breakpoint
    t["pos"] = np.where(swapped, t["p2"], t["p1"])
    t["end"] = np.where(swapped, t["p1"], t["p2"])
    # records near existing targets so that pairs survive both distance rules, plus pure noise
    pick = rng.integers(0, n_t, n_rec)
    near = rng.random(n_rec) < 0.7
    jitter = lambda: rng.integers(-2 * diff, 2 * diff + 1, n_rec)
    r = dict(g=np.where(near, t["g"][pick], rng.integers(0, n_pairs + 1, n_rec)),
             p1=np.clip(np.where(near, t["p1"][pick] + jitter(), rng.integers(0, span, n_rec)), 0, 0xFFFFFFFF),
             p2=np.clip(np.where(near, t["p2"][pick] + jitter(), rng.integers(0, span, n_rec)), 0, 0xFFFFFFFF))
    # duplicate keys: a quarter of the records repeat an earlier record exactly
    dup = np.flatnonzero(rng.random(n_rec) < 0.25)
    src = (dup * rng.random(dup.size)).astype(np.int64)
    for k in r:
        r[k][dup] = r[k][src]
    u32 = lambda d: {k: v.astype(np.uint32) for k, v in d.items()}
    return u32(t), u32(r)


def _tra_expected(port, t, r, diff):
    ql = np.clip(r["p1"].astype(np.int64) - diff, 0, 0xFFFFFFFF).astype(np.uint32)
    qh = np.clip(r["p1"].astype(np.int64) + diff, 0, 0xFFFFFFFF).astype(np.uint32)
    f = port.build(t["p1"], t["p1"], t["g"])
    off, tid = f.query_sorted_pairs(ql, qh, r["g"], threads=4)          # |p1 - p1'| <= diff and equal pair id
    rec = np.repeat(np.arange(ql.size), np.diff(off).astype(np.int64))
    i64 = lambda a: a.astype(np.int64)
    ok = np.abs(i64(r["p2"][rec]) - i64(t["p2"][tid])) <= diff
    n_pos, n_end = np.minimum(r["p1"], r["p2"])[rec], np.maximum(r["p1"], r["p2"])[rec]
    ok &= (n_pos <= t["end"][tid]) & (t["pos"][tid] <= n_end)            # find_overlaps on the raw intervals
    kept = [np.sort(tid[int(off[q]):int(off[q + 1])][ok[int(off[q]):int(off[q + 1])]]) for q in range(ql.size)]
    keys = np.stack([r["g"], r["g"] ^ 0x55, r["p1"], r["p2"]], axis=1).astype(np.uint32)
    return ql, qh, keys, kept


@pytest.mark.parametrize("seed,n_t,n_rec,probes", [(1, 3000, 2000, 1), (2, 3000, 2000, 3), (3, 1, 1, 1), (4, 50, 4000, 3),
                                                   (5, 20000, 30000, 3)])
def test_tra_rule_and_duplicate_keys(port_oracle, seed, n_t, n_rec, probes):
    diff, span = 5000, 400_000
    t, r = _tra_case(seed, n_t, n_rec, n_pairs=4, span=span, diff=diff)
    ql, qh, keys, kept = _tra_expected(port_oracle, t, r, diff)
    tra = dict(rec_p1=r["p1"], rec_p2=r["p2"], tgt_p1=t["p1"], tgt_p2=t["p2"], tgt_pos=t["pos"], tgt_end=t["end"])
    if probes == 1:
        tg, qg, q_lo, q_hi = t["g"], r["g"], ql, qh
    else:   # the C++ front end's keying: group = (pair, bucket of p2), one probe per bucket a partner can be in
        nb = 0xFFFFFFFF // diff + 1
        tg = (t["g"].astype(np.uint64) * nb + t["p2"] // diff).astype(np.uint32)
        mid = (r["p2"] // diff).astype(np.int64)
        qg = np.stack([np.where((mid + k >= 0) & (mid + k < nb), r["g"].astype(np.int64) * nb + mid + k, 0xFFFFFFFF)
                       for k in (-1, 0, 1)], axis=1).astype(np.uint32).reshape(-1)
        q_lo, q_hi = np.repeat(ql, 3), np.repeat(qh, 3)
    ix = DeviceIndex.build(t["p1"], t["p1"], tg)
    for dedup in (False, True):
        want = _dedup_first_with_hits(keys, kept) if dedup else kept
        off, tgt = ix.sv2nl_join(q_lo, q_hi, qg, diff=diff, probes_per_record=probes, tra=tra,
                                 rec_key=keys if dedup else None)
        assert off.size == n_rec + 1 and int(off[-1]) == tgt.size == sum(w.size for w in want)
        got = _per_record_sorted(off, tgt)
        assert all(np.array_equal(g, w) for g, w in zip(got, want))
    assert sum(k.size for k in kept) > 0 or n_t == 1
    # capacity too small: status + complete offsets/total, then the retry inside the binding succeeds
    off2, tgt2 = ix.sv2nl_join(q_lo, q_hi, qg, diff=diff, probes_per_record=probes, tra=tra, rec_key=keys, pair_capacity=1)
    assert np.array_equal(off2, off) and np.array_equal(np.sort(tgt2), np.sort(tgt))
    ix.close()


@pytest.mark.parametrize("kind", [_lib.FILTER_SV2NL_DUP, _lib.FILTER_SV2NL_INV])
def test_fused_filters_followed_by_the_duplicate_key_rule(port_oracle, kind):
    c = random_case(61, n_t=8000, n_q=6000, n_groups=4, span=300_000, max_len=20000, dup_frac=0.1)
    rng = np.random.default_rng(5)
    rep = np.flatnonzero(rng.random(c["ql"].size) < 0.3)                  # repeated NL records -> equal keys
    src = (rep * rng.random(rep.size)).astype(np.int64)
    for k in ("ql", "qh", "qg"):
        c[k][rep] = c[k][src]
    strand = rng.integers(0, 4, c["ql"].size).astype(np.uint8)
    strand[rep] = strand[src]
    keys = np.stack([c["qg"], np.full(c["ql"].size, 0xFFFFFFFF, np.uint32), c["ql"], c["qh"]], axis=1)
    ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
    off_f, _, ht_f = ix.join_filtered(c["ql"], c["qh"], c["qg"], kind=kind, diff=15000, use_strand=True, qstrand=strand)
    want = _dedup_first_with_hits(keys, _per_record_sorted(off_f, ht_f))  # (the filtered join is checked elsewhere)
    off, tgt = ix.sv2nl_join(c["ql"], c["qh"], c["qg"], kind=kind, diff=15000, use_strand=True, qstrand=strand,
                             rec_key=keys)
    got = _per_record_sorted(off, tgt)
    assert all(np.array_equal(g, w) for g, w in zip(got, want))
    assert 0 < int(off[-1]) < int(off_f[-1])                              # the rule removed something
    ix.close()


def test_bad_arguments_are_refused():
    ix = DeviceIndex.build(np.array([1, 5], np.uint32), np.array([3, 9], np.uint32))
    q = np.array([2, 4, 6], np.uint32)
    with pytest.raises(ValueError):
        ix.sv2nl_join(q, q, probes_per_record=2)                          # 3 queries, 2 per record
    with pytest.raises(ValueError):
        ix.sv2nl_join(q, q, rec_key=np.zeros((2, 4), np.uint32))           # one key per record
    rules = _lib.Sv2nlRules(0, 0, 0, 0, None, None, None, None, None, None, None)
    off = np.zeros(4, np.uint64)
    total = _lib.C.c_uint64()
    rc = _lib.load().bcu_sv2nl_join(ix._h, None, _lib.C.byref(rules), 3, None, q.ctypes.data, q.ctypes.data, None,
                                    off.ctypes.data, 0, None, _lib.C.byref(total))
    assert rc == _lib.BCU_E_INVALID and b"probes_per_record" in _lib.load().bcu_last_error()
    ix.close()
