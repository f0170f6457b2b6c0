"""Regenerates tests/golden/ref_*.json by running the UNMODIFIED reference IntervalTree
(oracle/_ref/libbinary_ref.so, built from /root/reference by oracle/Makefile). Run in the authoring
container only:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from cases import clrs_arrays, random_case  # noqa: E402


def dump(name, c, ref):
    f = ref.build(c["tl"], c["th"], c["tg"])
    off, tid, _ = f.query(c["ql"], c["qh"], c["qg"])
    groups = [0] if c["tg"] is None else sorted(set(map(int, c["tg"])))
    out = {k: (None if c[k] is None else list(map(int, c[k]))) for k in ("tl", "th", "tg", "ql", "qh", "qg")}
    out["offsets"] = list(map(int, off))
    out["targets_native"] = list(map(int, tid))
    out["roots"] = {str(g): f.root(g) for g in groups}
    out["black_height"] = {str(g): f.check_invariants(g) for g in groups}
    with open(os.path.join(HERE, name), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(name, "targets", len(out["tl"]), "queries", len(out["ql"]), "hits", out["offsets"][-1])


def main():
    ref = oracle.Oracle("reference")
    lo, hi = clrs_arrays()
    dump("ref_clrs.json", dict(tl=lo, th=hi, tg=None, ql=np.array([7, 15, 22, 100, 0, 26], np.uint32),
                               qh=np.array([25, 25, 25, 111, 0, 26], np.uint32), qg=None), ref)
    dump("ref_random_groups.json", random_case(101, n_t=400, n_q=150, span=20000, max_len=300, n_groups=4,
                                               q_groups=5), ref)
    dump("ref_edge_cases.json", random_case(102, n_t=300, n_q=120, span=5000, max_len=200, inverted_frac=0.2,
                                            dup_frac=0.2, extremes=True), ref)
    dump("ref_long_intervals.json", random_case(103, n_t=300, n_q=120, span=50000, max_len=100,
                                                long_frac=0.03, n_groups=2), ref)


if __name__ == "__main__":
    main()
