"""CPU: host-side logic of the N>1 path with world_size-2 gloo processes. The per-rank join is played by
the oracle here (no GPU in this container); tests/test_gpu_parity.py::test_sharded_join_on_gpu runs
the same logic over the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from binary_b200.sharding import ShardedJoin, assemble, broadcast_blob, replicate_index, shard_range
from cases import canonical, random_case


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _oracle_join(case):
    import oracle
    port = oracle.Oracle("port")
    forest = port.build(case["tl"], case["th"], case["tg"])

    def join(ql, qh, qg, qid_base):
        off, tid = forest.query_sorted_pairs(ql, qh, qg)
        hq = np.repeat(np.arange(ql.size, dtype=np.uint32), np.diff(off).astype(np.int64)) + np.uint32(qid_base)
        return off, hq, tid
    return join


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = random_case(77, n_t=4000, n_q=3001, n_groups=4, q_groups=5, dup_frac=0.1, inverted_frac=0.05)
    sj = ShardedJoin(_oracle_join(case), rank, world)
    piece = sj.run(case["ql"], case["qh"], case["qg"])
    piece["global_offsets"] = sj.global_offsets(piece)
    dist.barrier()
    np.savez(os.path.join(out_dir, f"piece{rank}.npz"), **piece)
    dist.destroy_process_group()


def test_two_rank_sharded_join_equals_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    pieces = []
    for r in range(2):
        z = np.load(tmp_path / f"piece{r}.npz")
        pieces.append({k: z[k] if z[k].ndim else int(z[k]) for k in z.files})
    off, hq, ht = assemble(pieces)
    case = random_case(77, n_t=4000, n_q=3001, n_groups=4, q_groups=5, dup_frac=0.1, inverted_frac=0.05)
    want_off, want_hq, want_ht = _oracle_join(case)(case["ql"], case["qh"], case["qg"], 0)
    assert np.array_equal(off, want_off)
    assert np.array_equal(hq, want_hq)            # global query ids, already in order
    assert np.array_equal(canonical(off, ht)[1], want_ht)
    # the 8-byte total exchange gives every rank its global offsets without moving any pairs
    assert np.array_equal(pieces[0]["global_offsets"], want_off[: pieces[0]["stop"] + 1])
    assert np.array_equal(pieces[1]["global_offsets"], want_off[pieces[1]["start"]:])


def _blob_worker(rank, world, port, out_dir):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src = 1  # not rank 0 on purpose
    payload = torch.arange(100_003, dtype=torch.int64).to(torch.uint8) if rank == src else None
    got = broadcast_blob(payload, src=src, device="cpu")
    empty = broadcast_blob(torch.empty(0, dtype=torch.uint8) if rank == src else None, src=src, device="cpu")
    np.save(os.path.join(out_dir, f"blob{rank}.npy"), got.numpy())
    assert empty.numel() == 0
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_blob_ships_size_then_payload(tmp_path):
    """Host logic of the build-once/broadcast path (replicate_index): receivers learn the size first."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_blob_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = (np.arange(100_003, dtype=np.int64) & 0xFF).astype(np.uint8)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"blob{r}.npy"), want)


def test_replicate_index_without_a_process_group_just_builds():
    marker = object()
    assert replicate_index(lambda: marker, device=0) is marker
