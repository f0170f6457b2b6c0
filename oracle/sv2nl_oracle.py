"""TEST INFRASTRUCTURE ONLY -- CPU restatement of sv2nl's mapping loop, record at a time.

Follows the reference line by line (paths relative to /root/reference/standalone/sv2nl):
  Mapper::map_delegate / map_impl / build_tree      include/mapper.hpp:147-162, 194-246
  Dup/Inv/TraMapper::check_condition, Tra build     source/mapper.cpp:50-79, 86-170
  validate_record, is_contained, distance_less,
  get_2chroms_with_pos, format_map_key               include/helper.hpp:16-91
  Writer::format_keys, header                        source/writer.cpp:21-27, include/mapper.hpp:30
  run(): the three mappers and their types           source/main.cpp:46-81
over the ORACLE interval tree (one tree per chromosome for Dup/Inv, one tree for Tra).

PARITY PINNED (round 2): the reference has no sv2nl tests or golden outputs of its own, but its sv2nl
sources compile here unmodified over a text-VCF stand-in for htslib (oracle/_ref/libsv2nl_ref.so, see
oracle/sv2nl_ref_harness.cpp and oracle/stubs/). tests/test_sv2nl_reference.py checks this restatement
against it predicate by predicate and as whole runs of the reference tool (fixture + synthetic VCFs, incl.
the strand carry-over of the reference reader); tests/golden/sv2nl_expected.json is the reference tool's own
output. What remains this repo's reading is only htslib's field extraction (oracle/stubs/htslib_text.cpp).
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, List

import numpy as np

HEADER = "chrom\tpos\tend\tsvtype\tchrom\tpos\tend\tsvtype"


@dataclass(frozen=True)
class Rec:
    chrom: str
    pos: int
    svend: int
    svtype: str
    chr2: str
    strand1: bool
    strand2: bool


def records(table) -> List[Rec]:
    return [Rec(str(table.chrom[i]), int(table.pos[i]), int(table.svend[i]), str(table.svtype[i]),
                str(table.chr2[i]), bool(table.strand1[i]), bool(table.strand2[i])) for i in range(len(table))]


def validate_record(r: Rec) -> Rec:  # helper.hpp:52-63
    if r.pos > r.svend:
        if r.svtype in ("BND", "TRA"):
            return replace(r, pos=r.svend, svend=r.pos, chrom=r.chr2, chr2=r.chrom)
        return replace(r, pos=r.svend, svend=r.pos)
    return r


def is_contained(target: Rec, source: Rec) -> bool:  # helper.hpp:16-25
    return target.pos <= source.pos and target.svend >= source.svend


def distance_less(a: Rec, b: Rec, threshold: int) -> bool:  # helper.hpp:32-41
    return abs(a.pos - b.pos) <= threshold and abs(a.svend - b.svend) <= threshold


def two_chroms_with_pos(r: Rec):  # helper.hpp:76-82
    return (r.chr2, r.svend, r.chrom, r.pos) if r.chrom > r.chr2 else (r.chrom, r.pos, r.chr2, r.svend)


def format_map_key(r: Rec) -> str:  # helper.hpp:84-91
    if r.svtype in ("TRA", "BND"):
        c1, p1, c2, p2 = two_chroms_with_pos(r)
        return f"{c1}-{c2}-{p1}-{p2}"
    return f"{r.chrom}-{r.pos}-{r.svend}"


def format_keys(r: Rec) -> str:  # writer.cpp:21-27 (note pos + 1)
    if r.svtype in ("TRA", "BND"):
        return f"{r.chrom},{r.chr2}\t{r.pos + 1}\t{r.svend}\t{r.svtype}"
    return f"{r.chrom}\t{r.pos + 1}\t{r.svend}\t{r.svtype}"


def check_dup(nl: Rec, sv: Rec, diff: int, use_strand: bool) -> bool:  # mapper.cpp:50-55
    return is_contained(sv, nl) and distance_less(nl, sv, diff)


def check_inv(nl: Rec, sv: Rec, diff: int, use_strand: bool) -> bool:  # mapper.cpp:57-79
    if is_contained(sv, nl) or is_contained(nl, sv) or not distance_less(nl, sv, diff):
        return False
    if not use_strand:
        return True
    if nl.pos <= sv.pos:
        return nl.strand1 and not nl.strand2
    return (not nl.strand1) and nl.strand2


def check_tra(nl: Rec, sv: Rec, diff: int, use_strand: bool) -> bool:  # mapper.cpp:144-156
    n1, np1, n2, np2 = two_chroms_with_pos(nl)
    s1, sp1, s2, sp2 = two_chroms_with_pos(sv)
    if n1 == s1 and n2 == s2:
        return abs(np1 - sp1) <= diff and abs(np2 - sp2) <= diff
    return False


def _tree(oracle, recs: List[Rec]):
    lo = np.array([r.pos for r in recs], dtype=np.uint32)
    hi = np.array([r.svend for r in recs], dtype=np.uint32)
    return oracle.build(lo, hi)


def _map(oracle, nl: List[Rec], chroms: List[str], trees: Dict[str, tuple], nl_type: str, check, diff: int,
         use_strand: bool) -> List[str]:
    lines: List[str] = []
    cache = set()  # SV2NL_USE_CACHE is ON (options.hpp:8): keys of NL records already written
    for chrom in chroms:  # one task per contig without '_' (mapper.hpp:239-244); order is irrelevant
        tree, tree_recs = trees[chrom]
        for r in nl:
            if r.chrom != chrom or r.svtype != nl_type:
                continue
            key = format_map_key(r)
            if key in cache:
                continue
            q = validate_record(r)
            _, hits, _ = tree.query([q.pos], [q.svend])
            kept = [tree_recs[t] for t in hits if check(q, tree_recs[t], diff, use_strand)]
            if kept:  # overlaps_vector.size() > 1
                cache.add(key)
                for sv in kept:
                    lines.append(format_keys(r) + "\t" + format_keys(sv))
    return lines


def sv2nl(oracle, nl_table, sv_table, diff: int = 1_000_000, use_strand: bool = True) -> Dict[str, List[str]]:
    """Returns {"dup": [...], "inv": [...], "tra": [...]}: the data lines of the three output files
    (each file also starts with HEADER). Line order inside a file is not defined by the reference (thread
    interleaving, tree preorder): compare as sorted lists."""
    nl, sv = records(nl_table), records(sv_table)
    chroms = [c for c in nl_table.contigs if "_" not in c]
    out = {}
    for name, nl_type, sv_type, check in (("dup", "TDUP", "DUP", check_dup), ("inv", "INV", "INV", check_inv)):
        trees = {}
        for c in chroms:  # build_tree: filter chrom & svtype, validate, insert (mapper.hpp:147-162)
            recs = [validate_record(r) for r in sv if r.chrom == c and r.svtype == sv_type]
            trees[c] = (_tree(oracle, recs), recs)
        out[name] = _map(oracle, nl, chroms, trees, nl_type, check, diff, use_strand)
    # TraMapper: ONE tree over all BND records, inserted WITHOUT validate_record (mapper.cpp:158-170)
    bnd = [r for r in sv if r.svtype == "BND"]
    shared = (_tree(oracle, bnd), bnd)
    out["tra"] = _map(oracle, nl, chroms, {c: shared for c in chroms}, "TRA", check_tra, diff, use_strand)
    return out
