"""GPU (-m gpu): ``bcu_join_multi`` -- one call, several GPUs (include/binary_cuda.h; the reference's counterpart is
one pool task per chromosome over shared trees, sv2nl mapper.hpp:238-246). On a one-GPU box the call runs with a
single range; with two or more devices the batch is split and every range is answered on its own device. Either
way the result must be the single-device CSR, bit for bit (offsets) and as the sorted pair set (targets)."""
import numpy as np
import pytest

from binary_b200 import DeviceIndex, join_multi
from cases import canonical, random_case

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 3, 8])
def test_join_multi_equals_oracle(port_oracle, n_dev, monkeypatch):
    if n_dev > _device_count():
        pytest.skip(f"needs {n_dev} GPUs")
    monkeypatch.setenv("BCU_HOST_CHUNK", "8192")     # several pipeline chunks per range (read once per process)
    c = random_case(70 + n_dev, n_t=40000, n_q=50021, n_groups=5, span=3_000_000, max_len=2500, dup_frac=0.03,
                    long_frac=0.001)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    counts = np.diff(want_off).astype(np.int64)
    want_q = np.repeat(np.arange(counts.size, dtype=np.uint32), counts)
    indexes = [DeviceIndex.build(c["tl"], c["th"], c["tg"], device=d) for d in range(n_dev)]
    off, hq, ht = join_multi(indexes, c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off, want_off) and np.array_equal(hq, want_q)
    assert np.array_equal(canonical(off, ht)[1], want_tid)
    # u32 counts instead of u64 offsets, no query-id column (the cheapest result over PCIe)
    cnt, hq2, ht2 = join_multi(indexes, c["ql"], c["qh"], c["qg"], want_query_ids=False, counts32=True)
    assert hq2 is None and np.array_equal(cnt.astype(np.int64), counts)
    assert np.array_equal(canonical(want_off, ht2)[1], want_tid)
    # a pair buffer that is too small is reported with the required size (the wrapper grows and retries)
    off3, _, ht3 = join_multi(indexes, c["ql"], c["qh"], c["qg"], pair_capacity=17)
    assert np.array_equal(off3, want_off) and np.array_equal(canonical(off3, ht3)[1], want_tid)
    # empty batch, batch smaller than the device count
    e = np.empty(0, np.uint32)
    off0, hq0, ht0 = join_multi(indexes, e, e, e)
    assert list(off0) == [0] and ht0.size == 0
    off1, _, ht1 = join_multi(indexes, c["ql"][:1], c["qh"][:1], c["qg"][:1])
    assert np.array_equal(off1, want_off[:2]) and np.array_equal(np.sort(ht1), want_tid[:int(want_off[1])])
    for ix in indexes:
        ix.close()


def test_python_drop_in_over_every_visible_gpu(port_oracle):
    """``IntervalTree(devices=[...])``: batches through bcu_join_multi over replicas, single queries on the first."""
    from binary_b200 import IntervalTree
    c = random_case(91, n_t=30000, n_q=40000, n_groups=3, span=2_000_000, max_len=3000)
    f = port_oracle.build(c["tl"], c["th"], c["tg"])
    want_off, want_tid = f.query_sorted_pairs(c["ql"], c["qh"], c["qg"], threads=4)
    t = IntervalTree(devices=list(range(_device_count())))
    t.insert_node(c["tl"], c["th"], c["tg"])
    off, tid = t.find_overlaps_batch(c["ql"], c["qh"], c["qg"])
    assert np.array_equal(off, want_off) and np.array_equal(canonical(off, tid)[1], want_tid)
    q = int(np.argmax(np.diff(want_off)))
    hits = t.find_overlaps(int(c["ql"][q]), int(c["qh"][q]), int(c["qg"][q]))
    assert sorted(h[2] for h in hits) == sorted(want_tid[int(want_off[q]):int(want_off[q + 1])].tolist())
    t.insert_node(5, 6, 0)                                      # replicas are rebuilt after an insert
    assert len(t.find_overlaps(5, 5, 0)) >= 1 and t.size() == 30001
