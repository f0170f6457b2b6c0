"""sv2nl's mapping loop on top of the batched GPU join.

Reference being restructured (``/root/reference/standalone/sv2nl``): ``Mapper::map_impl``
(``include/mapper.hpp:194-236``) builds one tree per chromosome and calls ``find_overlaps`` once per NL
record; ``TraMapper`` (``source/mapper.cpp:86-170``) shares one tree over all BND records. Here each of the
three mappers issues ONE batched join for all chromosomes (chromosome = ``group``; Tra joins on the
selective breakpoint-proximity condition instead of the raw interval, see below),
with the DUP / INV post-filters (``check_condition``, ``mapper.cpp:50-79``) fused into the join kernels
(``bcu_join_filtered``) and the TRA conditions (``mapper.cpp:144-156``) vectorised on the host, then the duplicate-key rule of ``SV2NL_USE_CACHE``
(``mapper.hpp:212-234``) and the writer's formatting (``writer.cpp:21-27``).

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


def _absdiff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a, b = a.astype(np.int64), b.astype(np.int64)
    return np.abs(a - b)


def _fmt(chrom, pos, end, svtype, chr2=None) -> str:  # writer.cpp:21-27, note pos + 1
    if chr2 is not None:
        return f"{chrom},{chr2}\t{int(pos) + 1}\t{int(end)}\t{svtype}"
    return f"{chrom}\t{int(pos) + 1}\t{int(end)}\t{svtype}"


def _first_per_key_with_hits(keys: List[str], has_hits: np.ndarray) -> np.ndarray:
    """SV2NL_USE_CACHE: an NL record is written only if no EARLIER record with the same key was written."""
    seen, keep = set(), np.zeros(len(keys), dtype=bool)
    for i, k in enumerate(keys):
        if has_hits[i] and k not in seen:
            seen.add(k)
            keep[i] = True
    return keep


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
            # check_condition is fused into the join kernels (bcu_join_filtered): rejected pairs are never
            # counted or written. The host re-evaluation below is then a no-op kept as a cross-check.
            qstrand = (nl.strand1[qsel].astype(np.uint8) | (nl.strand2[qsel].astype(np.uint8) << 1))
            off, hq, ht = ix.join_filtered(q_pos, q_end, cid(q_chrom),
                                           kind=_lib.FILTER_SV2NL_DUP if name == "dup" else _lib.FILTER_SV2NL_INV,
                                           diff=diff, use_strand=use_strand, qstrand=qstrand)
            ix.close()
            n_device = hq.size
            nlp, nle, svp, sve = q_pos[hq], q_end[hq], t_pos[ht], t_end[ht]
            sv_has_nl = (svp <= nlp) & (sve >= nle)                   # is_contained(sv, nl)
            near = (_absdiff(nlp, svp) <= diff) & (_absdiff(nle, sve) <= diff)  # distance_less
            if name == "dup":
                ok = sv_has_nl & near
            else:
                nl_has_sv = (nlp <= svp) & (nle >= sve)
                ok = ~sv_has_nl & ~nl_has_sv & near
                if use_strand:
                    s1, s2 = nl.strand1[qsel][hq], nl.strand2[qsel][hq]
                    left = nlp <= svp
                    ok &= np.where(left, s1 & ~s2, ~s1 & s2)
            hq, ht = hq[ok], ht[ok]
            if hq.size != n_device:
                raise AssertionError("device filter and host check_condition disagree")
            kept = np.bincount(hq, minlength=qsel.size) > 0
            keys = [f"{nl.chrom[i]}-{int(nl.pos[i])}-{int(nl.svend[i])}" for i in qsel]  # helper.hpp:84-91
            write = _first_per_key_with_hits(keys, kept)
            for q, t in zip(hq, ht):
                if write[q]:
                    i, k = qsel[q], tsel[t]
                    lines.append(_fmt(nl.chrom[i], nl.pos[i], nl.svend[i], nl.svtype[i]) + "\t" +
                                 _fmt(t_chrom[t], t_pos[t], t_end[t], sv.svtype[k]))
        out[name] = lines

    # TraMapper. The reference joins on the raw [pos, POS2] intervals of ALL BND records (one tree, not
    # validated, chromosome not part of the key) and then keeps the pairs with equal ordered chromosome
    # pairs and both breakpoints within `diff` (mapper.cpp:144-156). The raw interval of a translocation
    # spans two chromosomes' coordinates, so that join returns a large fraction of all BND records per query.
    # Same result, far fewer pairs: join on the SELECTIVE condition -- group = ordered chromosome pair,
    # target = the point p1, query = [p1 - diff, p1 + diff] -- and apply the remaining conditions (second
    # breakpoint within diff, and the reference's raw-interval overlap, which can still reject a pair) on
    # the host.
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
        off, hq, ht = ix.join(q_lo, q_hi, q_group)
        ix.close()
        t_pos, t_end = sv.pos[tsel], sv.svend[tsel]
        ok = (_absdiff(np1[hq], sp1[ht]) <= diff) & (_absdiff(np2[hq], sp2[ht]) <= diff)
        ok &= (q_pos[hq] <= t_end[ht]) & (t_pos[ht] <= q_end[hq])   # the reference's raw overlap, as written
        hq, ht = hq[ok], ht[ok]
        kept = np.bincount(hq, minlength=qsel.size) > 0
        keys = []
        for i in qsel:  # format_map_key of the ORIGINAL record
            c, c2, p, e = nl.chrom[i], nl.chr2[i], int(nl.pos[i]), int(nl.svend[i])
            keys.append(f"{c2}-{c}-{e}-{p}" if c > c2 else f"{c}-{c2}-{p}-{e}")
        write = _first_per_key_with_hits(keys, kept)
        for q, t in zip(hq, ht):
            if write[q]:
                i, k = qsel[q], tsel[t]
                lines.append(_fmt(nl.chrom[i], nl.pos[i], nl.svend[i], nl.svtype[i], nl.chr2[i]) + "\t" +
                             _fmt(sv.chrom[k], sv.pos[k], sv.svend[k], sv.svtype[k], sv.chr2[k]))
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
