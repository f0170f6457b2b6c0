"""Shared input builders for the parity tests (seeded; no reference files are read at run time)."""
import numpy as np

# The reference's own fixture: test/source/test_algorithm/test_interval_tree.cpp:88-92,112-116
CLRS_NODES = [(16, 21), (8, 9), (5, 8), (0, 3), (6, 10), (15, 23), (25, 30), (17, 19), (19, 20), (26, 26)]
# SURVEY.md 8(c) [probe]: native (preorder) hit order of the unmodified reference for these queries
CLRS_Q_7_25_NATIVE = [(16, 21), (8, 9), (5, 8), (6, 10), (15, 23), (19, 20), (17, 19), (25, 30)]
CLRS_Q_15_25_NATIVE = [(16, 21), (15, 23), (19, 20), (17, 19), (25, 30)]

U32_MAX = 0xFFFFFFFF


def clrs_arrays():
    lo = np.array([a for a, _ in CLRS_NODES], np.uint32)
    hi = np.array([b for _, b in CLRS_NODES], np.uint32)
    return lo, hi


def random_case(seed, n_t, n_q, span=1_000_000, max_len=2_000, n_groups=1, q_groups=None,
                inverted_frac=0.0, dup_frac=0.0, extremes=False, long_frac=0.0):
    """Random targets/queries with the edge cases the reference path admits: duplicates
    (test_interval_tree.cpp:146-155), low == 0, inverted low > high (TraMapper, mapper.cpp:127-142),
    coordinates at the u32 limits, a few very long intervals, unknown query groups."""
    rng = np.random.default_rng(seed)
    tl = rng.integers(0, span, n_t, dtype=np.uint64)
    th = tl + rng.integers(0, max_len + 1, n_t, dtype=np.uint64)
    if long_frac > 0 and n_t:
        k = max(1, int(n_t * long_frac))
        idx = rng.choice(n_t, k, replace=False)
        th[idx] = tl[idx] + rng.integers(span // 4, span, k, dtype=np.uint64)
    if inverted_frac > 0 and n_t:
        k = max(1, int(n_t * inverted_frac))
        idx = rng.choice(n_t, k, replace=False)
        th[idx] = tl[idx] // 2
    if dup_frac > 0 and n_t > 1:
        k = max(1, int(n_t * dup_frac))
        src = rng.integers(0, n_t, k)
        dst = rng.integers(0, n_t, k)
        tl[dst], th[dst] = tl[src], th[src]
    if extremes and n_t >= 8:
        tl[0], th[0] = 0, 0
        tl[1], th[1] = 0, U32_MAX
        tl[2], th[2] = U32_MAX, U32_MAX
        tl[3], th[3] = U32_MAX - 5, U32_MAX
        tl[4], th[4] = 0, 5
        tl[5], th[5] = U32_MAX, 0  # inverted across the whole range
    tl = np.minimum(tl, U32_MAX).astype(np.uint32)
    th = np.minimum(th, U32_MAX).astype(np.uint32)
    tg = rng.integers(0, n_groups, n_t).astype(np.uint32) if n_groups > 1 else None
    ql = rng.integers(0, span, n_q, dtype=np.uint64)
    qh = ql + rng.integers(0, max_len + 1, n_q, dtype=np.uint64)
    if extremes and n_q >= 8:
        ql[0], qh[0] = 0, 0
        ql[1], qh[1] = 0, U32_MAX
        ql[2], qh[2] = U32_MAX, U32_MAX
        ql[3], qh[3] = U32_MAX, 0  # inverted query
        ql[4], qh[4] = span * 2, span * 3  # beyond every target (but for the extremes)
        ql[5], qh[5] = 5, 5
    ql = np.minimum(ql, U32_MAX).astype(np.uint32)
    qh = np.minimum(qh, U32_MAX).astype(np.uint32)
    qgn = q_groups if q_groups is not None else n_groups
    qg = rng.integers(0, qgn, n_q).astype(np.uint32) if (n_groups > 1 or q_groups) else None
    if tg is not None and qg is None:
        qg = np.zeros(n_q, np.uint32)
    return dict(tl=tl, th=th, tg=tg, ql=ql, qh=qh, qg=qg)


def canonical(offsets, targets):
    """CSR -> canonical form: offsets + targets sorted ascending inside every query."""
    from oracle import sort_within_segments
    return np.asarray(offsets, np.uint64), sort_within_segments(np.asarray(offsets, np.uint64),
                                                                np.asarray(targets, np.uint32))
