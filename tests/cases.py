"""Shared input builders for the parity tests (seeded; no reference files are read at run time)."""
import os
from contextlib import contextmanager

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


# ---- synthetic sv2nl inputs (text VCFs with every record type the three mappers read) -----------------
SYNTH_CONTIGS = [("chr1", 2_000_000), ("chr2", 1_500_000), ("chr10", 1_200_000), ("chrX", 900_000),
                 ("chrUn_KI270302v1", 2274), ("chr1_KI270706v1_random", 175055)]


# htslib types INFO values by the header's ##INFO lines (an undeclared tag is a String: asking for END as an
# integer then fails), so every generated VCF declares what sv2nl reads (vcf_info.cpp:10-42)
VCF_INFO_HEADER = ['##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">', '##INFO=<ID=CHR2,Number=1,Type=String,Description="Chromosome for the second breakpoint">', '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the structural variant">', '##INFO=<ID=POS2,Number=1,Type=Integer,Description="Position of the second breakpoint (BND)">', '##INFO=<ID=SVEND,Number=1,Type=Integer,Description="2nd position of the structural variant">', '##INFO=<ID=STRAND1,Number=1,Type=String,Description="Strand for breakpoint1">', '##INFO=<ID=STRAND2,Number=1,Type=String,Description="Strand for breakpoint2">']


def write_synth_vcfs(directory, seed=1, n_sv=3000, n_nl=2000):
    """Delly-style SV VCF (DUP/INV/BND/DEL) + ScanNLS-style NL VCF (TDUP/INV/TRA/INS). NL records are
    jittered copies of SV records (so the filters see near hits, containments, strand cases, duplicates,
    POS > END swaps, records on '_' contigs) plus random noise. Returns (nl_path, sv_path)."""
    import os
    rng = np.random.default_rng(seed)
    names = [c for c, _ in SYNTH_CONTIGS]
    lens = dict(SYNTH_CONTIGS)
    head = ["##fileformat=VCFv4.2"] + [f"##contig=<ID={c},length={l}>" for c, l in SYNTH_CONTIGS]
    head += VCF_INFO_HEADER
    head.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO")
    sv_lines, nl_lines, sv_recs = [], [], []
    for i in range(n_sv):
        t = rng.choice(["DUP", "INV", "BND", "DEL"], p=[0.4, 0.25, 0.3, 0.05])
        c = names[rng.integers(0, 5)]
        L = int(rng.integers(1, 40_000))
        p = int(rng.integers(1, max(2, lens[c] - L)))
        if t == "BND":
            c2 = names[rng.integers(0, 4)]
            p2 = int(rng.integers(1, lens[c2]))
            sv_lines.append(f"{c}\t{p}\tsv{i}\tN\t<BND>\t.\tPASS\tSVTYPE=BND;CHR2={c2};POS2={p2}")
            sv_recs.append((t, c, p, c2, p2))
        else:
            e = p + L
            if rng.random() < 0.1:
                p, e = e, p  # POS > END: validate_record swaps
            sv_lines.append(f"{c}\t{p}\tsv{i}\tN\t<{t}>\t.\tPASS\tSVTYPE={t};END={e}")
            sv_recs.append((t, c, p, c, e))
    for i in range(n_nl):
        if rng.random() < 0.75:
            t, c, p, c2, e = sv_recs[rng.integers(0, n_sv)]
            j = lambda: int(rng.integers(-3000, 3000))
            p, e = max(1, p + j()), max(1, e + j())
            kind = {"DUP": "TDUP", "INV": "INV", "BND": "TRA", "DEL": "INS"}[t]
            if rng.random() < 0.15:
                p, e = e, p
                if kind == "TRA":
                    c, c2 = c2, c
        else:
            kind = rng.choice(["TDUP", "INV", "TRA", "INS"])
            c = names[rng.integers(0, 6)]
            c2 = names[rng.integers(0, 4)] if kind == "TRA" else c
            p = int(rng.integers(1, lens[c]))
            e = int(rng.integers(1, lens[c2]))
        s1, s2 = rng.choice(["+", "-"]), rng.choice(["+", "-"])
        info = f"SVTYPE={kind};CHR2={c2};SVEND={e}"
        if rng.random() < 0.9:
            info += f";STRAND1={s1}" + (f";STRAND2={s2}" if rng.random() < 0.95 else "")
        line = f"{c}\t{p}\tnl{i}\tN\t<{kind}>\t.\t.\t{info}"
        nl_lines.append(line)
        if rng.random() < 0.05:
            nl_lines.append(line)  # exact duplicate: SV2NL_USE_CACHE writes it once
    nl_path, sv_path = os.path.join(directory, "nl.vcf"), os.path.join(directory, "sv.vcf.gz")
    with open(nl_path, "w") as fh:
        fh.write("\n".join(head + nl_lines) + "\n")
    import gzip
    with gzip.open(sv_path, "wt") as fh:
        fh.write("\n".join(head + sv_lines) + "\n")
    return nl_path, sv_path


def write_config_e_vcfs(directory, n_sv=1_000_000, n_nl=5_000_000):
    """BASELINE.json configs[4] ("E"): a synthetic delly-style SV VCF (DUP/INV/BND) and a ScanNLS-style NL VCF
    (TDUP/INV/TRA) with the hg38 chromosome law and config B's length law, as VCF text. Returns (nl_path, sv_path)."""
    import os
    from binary_b200 import synth
    names = np.array(synth.HG38_NAMES, dtype=object)
    head = ["##fileformat=VCFv4.2"] + [f"##contig=<ID={n},length={l}>" for n, l in synth.HG38] + VCF_INFO_HEADER + \
           ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]

    def write(path, seed, n, kinds, probs, end_key, nls):
        g, lo, hi = synth.intervals_mt(seed, 0, n, "loguniform", 50, 10_000)
        rng = np.random.default_rng(seed)
        kind = rng.choice(np.array(kinds, dtype=object), size=n, p=probs)
        chr2 = names[rng.integers(0, 24, n)]
        pos2 = rng.integers(1, 40_000_000, n)
        s1 = rng.choice(np.array(["+", "-"], dtype=object), n)
        s2 = rng.choice(np.array(["+", "-"], dtype=object), n)
        with open(path, "w") as fh:
            fh.write("\n".join(head) + "\n")
            for b in range(0, n, 500_000):
                out = []
                for i in range(b, min(n, b + 500_000)):
                    k, c = kind[i], names[g[i]]
                    if k in ("BND", "TRA"):
                        info = f"SVTYPE={k};CHR2={chr2[i]};" + (f"POS2={pos2[i]}" if k == "BND" else f"SVEND={pos2[i]}")
                    else:
                        info = f"SVTYPE={k};{end_key}={hi[i] + 1}"
                        if nls:
                            info += f";STRAND1={s1[i]};STRAND2={s2[i]}"
                    out.append(f"{c}\t{lo[i] + 1}\tr{i}\tN\t<{k}>\t.\t.\t{info}")
                fh.write("\n".join(out) + "\n")

    sv_path, nl_path = os.path.join(directory, "sv.vcf"), os.path.join(directory, "nl.vcf")
    write(sv_path, 0xE5A0, n_sv, ["DUP", "INV", "BND"], [0.5, 0.25, 0.25], "END", False)
    write(nl_path, 0xE5A1, n_nl, ["TDUP", "INV", "TRA"], [0.5, 0.25, 0.25], "SVEND", True)
    return nl_path, sv_path


def line_set_digest(lines):
    """(count, order-independent 64-bit hash) of a list of text lines: sum of the first 8 bytes of blake2b(line)."""
    import hashlib
    h = 0
    for l in lines:
        h = (h + int.from_bytes(hashlib.blake2b(l.encode(), digest_size=8).digest(), "little")) & 0xFFFFFFFFFFFFFFFF
    return len(lines), h


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
