"""Synthetic workloads of BASELINE.json's configs.

Counter-based: element ``i`` of a stream is a pure function ``splitmix64(seed, i)`` so host numpy, a
C++ host or a device kernel can regenerate any slice without state (used by the multi-GPU bench to
give every rank its own query shard without materialising the whole stream).

Chromosome law ("hg38"): the 25 contigs without ``_`` of the reference fixture header
(``test/data/debug_uncom.vcf`` lines 5-459), in HEADER order, which is also the reference's task
order (sv2nl mapper.hpp:239-244). A record's chromosome is drawn proportionally to contig length and
its start uniformly in ``[0, len_c - L]``; ``group`` = index in header order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

HG38 = (
    ("chr1", 248956422), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
    ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
    ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr2", 242193529),
    ("chr20", 64444167), ("chr21", 46709983), ("chr22", 50818468), ("chr3", 198295559),
    ("chr4", 190214555), ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973),
    ("chr8", 145138636), ("chr9", 138394717), ("chrM", 16569), ("chrX", 156040895),
    ("chrY", 57227415),
)
HG38_NAMES = tuple(n for n, _ in HG38)
HG38_LENGTHS = np.array([l for _, l in HG38], dtype=np.uint64)
HG38_TOTAL = int(HG38_LENGTHS.sum())  # 3,088,286,401
_CUM = np.concatenate(([0], np.cumsum(HG38_LENGTHS))).astype(np.uint64)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, idx: np.ndarray, stream: int = 0) -> np.ndarray:
    """``mix(seed + (3*idx + stream + 1) * golden)``: three independent u64 streams per element."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (np.uint64(3) * idx.astype(np.uint64) + np.uint64(stream + 1))
             * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _unit(u: np.ndarray) -> np.ndarray:
    """u64 -> float64 in [0,1) from the top 53 bits."""
    return (u >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def intervals(seed: int, start: int, count: int, len_law: str, len_lo: int, len_hi: int
              ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Elements ``[start, start+count)`` of stream ``seed`` -> ``(group, low, high)`` u32 arrays.

    ``len_law``: ``"loguniform"`` or ``"uniform"`` over ``[len_lo, len_hi]`` base pairs; the interval is
    CLOSED: ``high = low + L - 1`` (L >= 1).
    """
    idx = np.arange(start, start + count, dtype=np.uint64)
    u_chr, u_len, u_pos = (splitmix64(seed, idx, s) for s in range(3))
    # chromosome ~ length: a uniform base-pair position over the concatenated genome picks it
    gpos = (_unit(u_chr) * HG38_TOTAL).astype(np.uint64)
    group = (np.searchsorted(_CUM, gpos, side="right") - 1).astype(np.int64)
    group = np.clip(group, 0, len(HG38) - 1)
    clen = HG38_LENGTHS[group]
    if len_law == "loguniform":
        L = np.exp(np.log(len_lo) + _unit(u_len) * (np.log(len_hi + 1) - np.log(len_lo)))
    elif len_law == "uniform":
        L = len_lo + _unit(u_len) * (len_hi + 1 - len_lo)
    else:
        raise ValueError(len_law)
    L = np.clip(np.floor(L).astype(np.uint64), max(len_lo, 1), len_hi)
    L = np.minimum(L, clen)
    low = (_unit(u_pos) * (clen - L + np.uint64(1)).astype(np.float64)).astype(np.uint64)
    low = np.minimum(low, clen - L)
    high = low + L - np.uint64(1)
    return group.astype(np.uint32), low.astype(np.uint32), high.astype(np.uint32)


def intervals_mt(seed: int, start: int, count: int, len_law: str, len_lo: int, len_hi: int,
                 chunk: int = 1 << 22) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """:func:`intervals` over ``chunk``-sized pieces on a thread pool (the stream is counter-based, so the
    pieces are independent; numpy releases the GIL inside its loops). Same values, several times faster
    for the 100 M-query batch of config D."""
    if count <= chunk:
        return intervals(seed, start, count, len_law, len_lo, len_hi)
    import os
    from concurrent.futures import ThreadPoolExecutor
    group, low, high = (np.empty(count, np.uint32) for _ in range(3))

    def piece(off):
        n = min(chunk, count - off)
        g, l, h = intervals(seed, start + off, n, len_law, len_lo, len_hi)
        group[off:off + n], low[off:off + n], high[off:off + n] = g, l, h

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        list(pool.map(piece, range(0, count, chunk)))
    return group, low, high


@dataclass(frozen=True)
class Workload:
    name: str
    n_targets: int
    n_queries: int
    t_seed: int
    q_seed: int
    t_law: str
    t_len: Tuple[int, int]
    q_law: str = "loguniform"
    q_len: Tuple[int, int] = (1, 1000)

    def targets(self, count: Optional[int] = None):
        return intervals_mt(self.t_seed, 0, self.n_targets if count is None else count, self.t_law, *self.t_len)

    def queries(self, start: int = 0, count: Optional[int] = None):
        count = self.n_queries - start if count is None else count
        return intervals_mt(self.q_seed, start, count, self.q_law, *self.q_len)

    def scaled(self, n_targets: int, n_queries: int) -> "Workload":
        """Same laws and seeds at a reduced size (parity tests at oracle-friendly sizes)."""
        return Workload(f"{self.name}[{n_targets}x{n_queries}]", n_targets, n_queries, self.t_seed,
                        self.q_seed, self.t_law, self.t_len, self.q_law, self.q_len)


# B sparse, C dense/output-heavy, D scale-out
CONFIG_B = Workload("B:sparse-1Mx10M", 1_000_000, 10_000_000, 0xB1A0, 0xB1A1, "loguniform", (50, 10_000))
CONFIG_C = Workload("C:dense-1Mx10M", 1_000_000, 10_000_000, 0xB1A2, 0xB1A3, "uniform", (10_000, 500_000))
CONFIG_D = Workload("D:scaleout-10Mx100M", 10_000_000, 100_000_000, 0xB1A4, 0xB1A5, "loguniform", (50, 10_000))
CONFIGS = {"B": CONFIG_B, "C": CONFIG_C, "D": CONFIG_D}


def algorithmic_bytes(n_q: int, n_t: int, n_hits: int) -> int:
    """Query-phase algorithmic bytes: read q.low,q.high (8 B/query), write the u64 CSR
    offset (8 B/query), write (u32 query_id, u32 target_id) per hit (8 B/hit), read start,end,id once
    (12 B/target). Group ids, the directory and any re-reads are NOT credited."""
    return 8 * n_q + 8 * n_q + 8 * n_hits + 12 * n_t


def algorithmic_build_bytes(n_t: int, n_groups: int = 25) -> int:
    """Build-phase algorithmic bytes: P radix passes reading and writing a 16-byte record
    (u64 key group<<32|start, u32 end, u32 id), P = ceil((32 + ceil(log2 n_groups)) / 8), plus the index pass
    (read start,end 8 B, write max-end 4 B): ``P*2*16*n_t + 12*n_t``."""
    import math
    p = math.ceil((32 + math.ceil(math.log2(max(n_groups, 2)))) / 8)
    return p * 2 * 16 * n_t + 12 * n_t
