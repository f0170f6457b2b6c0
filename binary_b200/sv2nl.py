"""sv2nl's mapping loop on top of the batched GPU join.

Reference being restructured (``/root/reference/standalone/sv2nl``): ``Mapper::map_impl``
(``include/mapper.hpp:194-236``) builds one tree per chromosome and calls ``find_overlaps`` once per NL
record; ``TraMapper`` (``source/mapper.cpp:86-170``) shares one tree over all BND records. Here each of the
three mappers issues ONE batched join for all chromosomes (chromosome = ``group``; Tra joins on the
selective breakpoint-proximity condition instead of the raw interval, see below),
through ``bcu_sv2nl_join``: the DUP / INV post-filters (``check_condition``, ``mapper.cpp:50-79``) are fused into
the join kernels, the TRA conditions (``mapper.cpp:144-156``) and the duplicate-key rule of ``SV2NL_USE_CACHE``
(``mapper.hpp:212-234``) run on the device on the join's CSR (``csrc/sv2nl_rules.cu``); the host only formats the
surviving lines (``writer.cpp:21-27``).

Both VCFs are parsed once (``vcf_text.read_vcf``) instead of once per chromosome task. Output lines are
the same multiset as the reference's; their order is not defined there (thread interleaving).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import _lib
from .interval_tree import DeviceIndex
from .vcf_text import VcfTable, read_vcf

HEADER = "chrom\tpos\tend\tsvtype\tchrom\tpos\tend\tsvtype"


def _chrom_ids(*tables: VcfTable) -> Dict[str, int]:
    ids: Dict[str, int] = {}
    for t in tables:
        for name in list(t.contigs) + list(t.chrom) + list(t.chr2):
            if name not in ids:
                ids[name] = len(ids)
    return ids


def _validated(t: VcfTable, sel: np.ndarray, swap_chroms: bool):
    """validate_record (helper.hpp:52-63) on the selected rows: swap pos/svend (and chrom/chr2 for
    TRA/BND) where pos > svend. Returns (chrom, pos, svend, chr2) arrays."""
    chrom, pos, end, chr2 = t.chrom[sel].copy(), t.pos[sel].copy(), t.svend[sel].copy(), t.chr2[sel].copy()
    inv = pos > end
    pos[inv], end[inv] = end[inv], pos[inv].copy()
    if swap_chroms:
        chrom[inv], chr2[inv] = chr2[inv], chrom[inv].copy()
    return chrom, pos, end, chr2


def _fmt(chrom, pos, end, svtype, chr2=None) -> str:  # writer.cpp:21-27, note pos + 1
    if chr2 is not None:
        return f"{chrom},{chr2}\t{int(pos) + 1}\t{int(end)}\t{svtype}"
    return f"{chrom}\t{int(pos) + 1}\t{int(end)}\t{svtype}"


def map_sv2nl(nl: VcfTable, sv: VcfTable, diff: int = 1_000_000, use_strand: bool = True, device: int = 0
              ) -> Dict[str, List[str]]:
    """{"dup": [...], "inv": [...], "tra": [...]}: data lines of sv2nl's three output files."""
    ids = _chrom_ids(nl, sv)
    main = np.array(["_" not in c for c in nl.chrom], dtype=bool) & np.isin(nl.chrom, [c for c in nl.contigs])
    cid = lambda names: np.array([ids[n] for n in names], dtype=np.uint32)
    out: Dict[str, List[str]] = {}

    for name, nl_type, sv_type in (("dup", "TDUP", "DUP"), ("inv", "INV", "INV")):
        tsel = np.flatnonzero(sv.svtype == sv_type)
        qsel = np.flatnonzero((nl.svtype == nl_type) & main)
        lines: List[str] = []
        if tsel.size and qsel.size:
            t_chrom, t_pos, t_end, _ = _validated(sv, tsel, swap_chroms=False)  # build_tree validates
            q_chrom, q_pos, q_end, _ = _validated(nl, qsel, swap_chroms=False)
            ix = DeviceIndex.build(t_pos, t_end, cid(t_chrom), device=device)
            # check_condition is fused into the join kernels and the duplicate-key rule runs on the join's CSR on the
            # device: what comes back are the pairs that are written
            qstrand = (nl.strand1[qsel].astype(np.uint8) | (nl.strand2[qsel].astype(np.uint8) << 1))
            key = np.stack([cid(nl.chrom[qsel]), np.full(qsel.size, 0xFFFFFFFF, np.uint32),     # helper.hpp:84-91:
                            nl.pos[qsel].astype(np.uint32), nl.svend[qsel].astype(np.uint32)], axis=1)  # ORIGINAL record
            off, ht = ix.sv2nl_join(q_pos, q_end, cid(q_chrom),
                                    kind=_lib.FILTER_SV2NL_DUP if name == "dup" else _lib.FILTER_SV2NL_INV,
                                    diff=diff, use_strand=use_strand, qstrand=qstrand, rec_key=key)
            ix.close()
            for q in np.flatnonzero(np.diff(off)):
                i = qsel[q]
                left = _fmt(nl.chrom[i], nl.pos[i], nl.svend[i], nl.svtype[i])
                for t in ht[int(off[q]):int(off[q + 1])]:
                    lines.append(left + "\t" + _fmt(t_chrom[t], t_pos[t], t_end[t], sv.svtype[tsel[t]]))
        out[name] = lines

    # TraMapper. The reference joins on the raw [pos, POS2] intervals of ALL BND records (one tree, not
    # validated, chromosome not part of the key) and then keeps the pairs with equal ordered chromosome
    # pairs and both breakpoints within `diff` (mapper.cpp:144-156). The raw interval of a translocation
    # spans two chromosomes' coordinates, so that join returns a large fraction of all BND records per query.
    # Same result, far fewer pairs: join on the SELECTIVE condition -- group = ordered chromosome pair,
    # target = the point p1, query = [p1 - diff, p1 + diff] -- and apply the remaining conditions (second
    # breakpoint within diff, and the reference's raw-interval overlap, which can still reject a pair) on
    # the device as well (bcu_sv2nl_rules.tra).
    tsel = np.flatnonzero(sv.svtype == "BND")
    qsel = np.flatnonzero((nl.svtype == "TRA") & main)
    lines = []
    if tsel.size and qsel.size:
        q_chrom, q_pos, q_end, q_chr2 = _validated(nl, qsel, swap_chroms=True)

        def ordered(chrom, pos, chr2, end):  # get_2chroms_with_pos (helper.hpp:76-82)
            sw = (chrom > chr2).astype(bool)
            return (np.where(sw, chr2, chrom), np.where(sw, end, pos), np.where(sw, chrom, chr2),
                    np.where(sw, pos, end))
        n1, np1, n2, np2 = ordered(q_chrom, q_pos, q_chr2, q_end)
        s1, sp1, s2, sp2 = ordered(sv.chrom[tsel], sv.pos[tsel], sv.chr2[tsel], sv.svend[tsel])
        pair_ids: Dict[tuple, int] = {}
        pid = lambda a, b: np.array([pair_ids.setdefault((x, y), len(pair_ids)) for x, y in zip(a, b)],
                                    dtype=np.uint32)
        t_group, q_group = pid(s1, s2), pid(n1, n2)
        p1 = np1.astype(np.int64)
        q_lo = np.clip(p1 - diff, 0, 0xFFFFFFFF).astype(np.uint32)
        q_hi = np.clip(p1 + diff, 0, 0xFFFFFFFF).astype(np.uint32)
        sp1u = sp1.astype(np.uint32)
        ix = DeviceIndex.build(sp1u, sp1u, t_group, device=device)
        # format_map_key of the ORIGINAL record (helper.hpp:84-91): ordered chromosome pair + positions
        o_c, o_c2 = nl.chrom[qsel], nl.chr2[qsel]
        o_p, o_e = nl.pos[qsel].astype(np.uint32), nl.svend[qsel].astype(np.uint32)
        sw = o_c > o_c2
        key = np.stack([np.where(sw, cid(o_c2), cid(o_c)), np.where(sw, cid(o_c), cid(o_c2)),
                        np.where(sw, o_e, o_p), np.where(sw, o_p, o_e)], axis=1)
        tra = dict(rec_p1=np1, rec_p2=np2, tgt_p1=sp1, tgt_p2=sp2, tgt_pos=sv.pos[tsel], tgt_end=sv.svend[tsel])
        off, ht = ix.sv2nl_join(q_lo, q_hi, q_group, diff=diff, tra=tra, rec_key=key)
        ix.close()
        for q in np.flatnonzero(np.diff(off)):
            i = qsel[q]
            left = _fmt(nl.chrom[i], nl.pos[i], nl.svend[i], nl.svtype[i], nl.chr2[i])
            for t in ht[int(off[q]):int(off[q + 1])]:
                k = tsel[t]
                lines.append(left + "\t" + _fmt(sv.chrom[k], sv.pos[k], sv.svend[k], sv.svtype[k], sv.chr2[k]))
    out["tra"] = lines
    return out


def run(nl_path: str, sv_path: str, output: str, diff: int = 1_000_000, use_strand: bool = True,
        device: int = 0) -> Dict[str, int]:
    """sv2nl's ``run`` (source/main.cpp:46-81): writes ``<output>.dup/.inv/.tra``; returns line counts."""
    res = map_sv2nl(read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff, use_strand, device)
    for ext, lines in res.items():
        with open(f"{output}.{ext}", "w") as fh:
            fh.write(HEADER + "\n")
            for line in lines:
                fh.write(line + "\n")
    return {k: len(v) for k, v in res.items()}
