"""GPU (-m gpu): the code paths the bench headlines, at the sizes it headlines them on.

* the host-buffer pipeline of ``bcu_join`` / ``bcu_join_filtered`` (host_join.cu: chunks flowing through
  copy-in / run / copy-out streams, chained offset bases, segmented pair copies) with enough queries for
  several pipeline chunks, and once more with ``BCU_HOST_CHUNK=4096`` so that quarter/half/tail chunks and
  dozens of chunk boundaries are hit;
* BASELINE.json configs C (full 10 M queries) and D (10 M targets, a 16 M-query slice of the 100 M batch) by
  total, per-query counts and the order-independent pair hash against the CPU flat-index twin (itself pinned
  to the reference tree walk in tests/test_oracle.py).

Integer work: bit-exact, no tolerance anywhere. The reference semantics are find_overlaps_impl,
library/include/binary/algorithm/interval_tree.hpp:306-334.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from binary_b200 import DeviceIndex, synth
from cases import canonical, random_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _qid_column(offsets):
    counts = np.diff(offsets).astype(np.int64)
    return np.repeat(np.arange(counts.size, dtype=np.uint32), counts)


def test_host_pipeline_many_chunks_vs_flat_oracle(port_oracle):
    """bcu_join with HOST buffers on 1 M targets x 5 M queries of config B's law (>= 4 pipeline chunks of
    2 Mi queries incl. the quarter/half head and tail): offsets, per-query counts and pair hash."""
    w = synth.CONFIG_B
    n_q = 5_000_000
    tg, tl, th = w.targets()
    qg, ql, qh = w.queries(0, n_q)
    want_total, want_hash, want_counts = port_oracle.flat_count_hash(tl, th, ql, qh, tg, qg, want_counts=True)
    ix = DeviceIndex.build(tl, th, tg)
    off, hq, ht = ix.join(ql, qh, qg, pair_capacity=want_total + 7)
    assert int(off[-1]) == want_total and np.array_equal(np.diff(off), want_counts)
    assert np.array_equal(hq, _qid_column(off))
    assert port_oracle.pair_hash(hq, ht) == want_hash
    # hit_query = NULL variant (what bench.py's e2e leg calls): same offsets and targets
    off2, hq2, ht2 = ix.join(ql, qh, qg, pair_capacity=want_total, want_query_ids=False)
    assert hq2 is None and np.array_equal(off2, off)
    assert port_oracle.pair_hash(_qid_column(off2), ht2) == want_hash
    # u32 counts instead of u64 offsets (bcu_join_multi on one device = the same chunk pipeline): what bench.py's
    # e2e leg calls since round 2
    from binary_b200 import join_multi
    cnt3, hq3, ht3 = join_multi([ix], ql, qh, qg, pair_capacity=want_total, want_query_ids=False, counts32=True)
    assert hq3 is None and cnt3.dtype == np.uint32 and np.array_equal(cnt3, want_counts)
    assert port_oracle.pair_hash(_qid_column(off), ht3) == want_hash
    cnt4, hq4, ht4 = join_multi([ix], ql[:4097], qh[:4097], qg[:4097], counts32=True)      # with the query column
    assert np.array_equal(cnt4, want_counts[:4097]) and np.array_equal(hq4, _qid_column(off[:4098]))
    ix.close()


def _host_filter(kind, diff, use_strand, ql, qh, tl, th, strand):
    """numpy twin of sv2nl's check_condition (standalone/sv2nl/source/mapper.cpp:50-79) over pair columns."""
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


@pytest.mark.parametrize("kind,use_strand", [(1, False), (2, True)])
def test_host_pipeline_filtered_many_chunks(port_oracle, kind, use_strand):
    """bcu_join_filtered with host buffers over >= 4 pipeline chunks == the unfiltered join's pairs filtered
    with check_condition on the host (the unfiltered join is checked against the oracle above)."""
    w = synth.CONFIG_B
    n_q = 3_500_000
    tg, tl, th = w.targets()
    qg, ql, qh = w.queries(0, n_q)
    strand = (np.arange(n_q, dtype=np.uint64) * 2654435761 >> 7).astype(np.uint8) & 3
    diff = 4000
    ix = DeviceIndex.build(tl, th, tg)
    off, hq, ht = ix.join(ql, qh, qg)
    want_total, want_hash = port_oracle.flat_count_hash(tl, th, ql, qh, tg, qg)
    assert int(off[-1]) == want_total and port_oracle.pair_hash(hq, ht) == want_hash
    keep = _host_filter(kind, diff, use_strand, ql[hq], qh[hq], tl[ht], th[ht], strand[hq])
    want_q, want_t = hq[keep], ht[keep]
    want_off = np.zeros(n_q + 1, np.uint64)
    np.cumsum(np.bincount(want_q, minlength=n_q), out=want_off[1:])
    goff, ghq, ght = ix.join_filtered(ql, qh, qg, kind=kind, diff=diff, use_strand=use_strand, qstrand=strand)
    assert 0 < want_q.size < hq.size
    assert np.array_equal(goff, want_off) and np.array_equal(ghq, want_q)
    assert port_oracle.pair_hash(ghq, ght) == port_oracle.pair_hash(want_q, want_t)
    ix.close()


_SMALL_CHUNK_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np
import oracle
from binary_b200 import DeviceIndex
from cases import canonical, random_case
port = oracle.Oracle("port")
# mixed batch: short candidate ranges interleaved with long ones, several groups, duplicates, inverted rows
c = random_case(9, n_t=20000, n_q=30000, span=200000, max_len=60000, n_groups=3, dup_frac=0.05, inverted_frac=0.02)
s = random_case(10, n_t=20000, n_q=30000, span=200000, max_len=30, n_groups=3)
ql = np.stack([c["ql"], s["ql"]], axis=1).reshape(-1).copy()
qh = np.stack([c["qh"], s["qh"]], axis=1).reshape(-1).copy()
qg = np.stack([c["qg"], s["qg"]], axis=1).reshape(-1).copy()
f = port.build(c["tl"], c["th"], c["tg"])
want_off, want_tid = f.query_sorted_pairs(ql, qh, qg, threads=4)
ix = DeviceIndex.build(c["tl"], c["th"], c["tg"])
for n in (ql.size, 4096, 4097, 8191, 12289, 1):           # chunk-boundary sizes: exact multiples, +-1, tails
    off, hq, ht = ix.join(ql[:n], qh[:n], qg[:n])
    assert np.array_equal(off, want_off[:n + 1]), n
    assert np.array_equal(hq, np.repeat(np.arange(n, dtype=np.uint32), np.diff(off).astype(np.int64))), n
    assert np.array_equal(canonical(off, ht)[1], want_tid[:int(want_off[n])]), n
strand = (np.arange(ql.size) % 4).astype(np.uint8)
goff, ghq, ght = ix.join_filtered(ql, qh, qg, kind=2, diff=30000, use_strand=True, qstrand=strand)
off, hq, ht = ix.join(ql, qh, qg)
a, b, tl_, th_ = ql[hq].astype(np.int64), qh[hq].astype(np.int64), c["tl"][ht].astype(np.int64), c["th"][ht].astype(np.int64)
t_has_q = (tl_ <= a) & (th_ >= b); q_has_t = (a <= tl_) & (b >= th_)
near = (np.abs(a - tl_) <= 30000) & (np.abs(b - th_) <= 30000)
s1, s2 = (strand[hq] & 1).astype(bool), (strand[hq] & 2).astype(bool)
keep = ~t_has_q & ~q_has_t & near & np.where(a <= tl_, s1 & ~s2, ~s1 & s2)
want = np.zeros(ql.size + 1, np.uint64); np.cumsum(np.bincount(hq[keep], minlength=ql.size), out=want[1:])
assert keep.sum() > 0 and np.array_equal(goff, want) and np.array_equal(ghq, hq[keep])
assert np.array_equal(canonical(goff, ght)[1], canonical(want, ht[keep])[1])
print("SMALL_CHUNK_OK", ql.size, int(want_off[-1]))
"""


def test_host_pipeline_with_tiny_chunks_in_a_subprocess():
    """BCU_HOST_CHUNK is read once per process: run the mixed short/long batch with 4096-query chunks (15
    chunks incl. 1024/2048 head and tail pieces) against the TREE oracle, plain and filtered."""
    env = dict(os.environ, BCU_HOST_CHUNK="4096")
    script = _SMALL_CHUNK_SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SMALL_CHUNK_OK" in r.stdout, r.stdout + r.stderr


def _device_join_count_hash(port_oracle, w, n_t, q_start, n_q):
    import torch
    tg, tl, th = w.targets(n_t)
    qg, ql, qh = w.queries(q_start, n_q)
    want_total, want_hash, want_counts = port_oracle.flat_count_hash(tl, th, ql, qh, tg, qg, qid_base=q_start,
                                                                     want_counts=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
    d_tg, d_tl, d_th, d_qg, d_ql, d_qh = map(t, (tg, tl, th, qg, ql, qh))
    ix = DeviceIndex.build_dev(tl.size, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr())
    d_off = torch.empty(n_q + 1, dtype=torch.int64, device=dev)
    cap = want_total + 16
    d_hq = torch.full((cap,), -1, dtype=torch.int32, device=dev)
    d_ht = torch.full((cap,), -1, dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ix.join_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(), d_ht.data_ptr(),
                d_total.data_ptr(), d_qg.data_ptr(), q_start, stream)
    torch.cuda.synchronize()
    assert int(d_total.item()) == want_total
    off = d_off.cpu().numpy().view(np.uint64)
    assert off[0] == 0 and np.array_equal(np.diff(off), want_counts)
    hq = d_hq[:want_total].cpu().numpy().view(np.uint32)
    ht = d_ht[:want_total].cpu().numpy().view(np.uint32)
    assert np.array_equal(hq, _qid_column(off) + np.uint32(q_start))         # pairs sorted by query id
    assert port_oracle.pair_hash(hq, ht) == want_hash
    assert (d_hq[want_total:] == -1).all() and (d_ht[want_total:] == -1).all()   # nothing past the total
    # any-overlap bit on the same batch
    d_any = torch.empty(n_q, dtype=torch.uint8, device=dev)
    ix.any_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_any.data_ptr(), d_qg.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_any.cpu().numpy().astype(bool), want_counts > 0)
    info = ix.info()
    ix.close()
    return want_total, info


def test_config_c_full_size(port_oracle):
    """BASELINE.json configs[2], all 10 M queries (8.3e8 pairs): count + per-query counts + pair hash."""
    total, _ = _device_join_count_hash(port_oracle, synth.CONFIG_C, synth.CONFIG_C.n_targets, 0, 10_000_000)
    assert total > 8 * 10**8


def test_config_d_ten_million_targets(port_oracle):
    """BASELINE.json configs[3]: the 10 M-target index (larger than L2, several length classes) against a
    16 M-query slice taken from the MIDDLE of the 100 M-query stream (query_id_base != 0)."""
    total, info = _device_join_count_hash(port_oracle, synth.CONFIG_D, synth.CONFIG_D.n_targets, 37_500_000,
                                          16_000_000)
    assert total > 10**8 and info["n_targets"] == 10_000_000
