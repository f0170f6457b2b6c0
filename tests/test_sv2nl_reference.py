"""The sv2nl level pinned against the UNMODIFIED reference sources.

``oracle/_ref/libsv2nl_ref.so`` is the reference's own ``standalone/sv2nl/source/{mapper,vcf_info,writer}.cpp``
+ ``include/*.hpp`` + ``library/include/binary/parser/vcf.hpp`` compiled in the authoring container over a
text-VCF stand-in for htslib (``oracle/stubs/``; recipe: ``oracle/Makefile``). Against it are checked

* the three ``check_condition`` predicates, ``validate_record``, ``format_map_key`` and ``Writer::format_keys``
  of the CPU restatement (``oracle/sv2nl_oracle.py``) on random and boundary records (CPU);
* the whole mapping (readers incl. the strand carry-over of vcf_info.cpp:17-31, the per-chromosome trees, the
  duplicate-key cache rule of mapper.hpp:212-234, the writer) on the reduced reference fixture and on
  synthetic VCFs, for the restatement (CPU), the batched GPU mapping ``binary_b200/sv2nl.py`` and the C++ tool
  ``standalone/sv2nl`` (``-m gpu``);
* the device-side pair filters (``accept<>`` in join.cu through ``bcu_join_filtered``) pair by pair (``-m gpu``).
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import sv2nl_oracle, sv2nl_ref
from oracle.sv2nl_oracle import Rec
from binary_b200.vcf_text import read_vcf
from cases import write_synth_vcfs

pytestmark = pytest.mark.skipif(not sv2nl_ref.available(),
                                reason="oracle/_ref/libsv2nl_ref.so not built (needs /root/reference at build time)")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NL, SV = os.path.join(GOLDEN, "nl_fixture.vcf"), os.path.join(GOLDEN, "sv_fixture.vcf")
CHROMS = ["chr1", "chr10", "chr2", "chrX", "chr17"]


def _random_pairs(seed, n, kind):
    """Record pairs around each other at the scale of `diff` so that every branch of the predicates is hit:
    containment either way, equal ends, distances at diff and diff + 1, inverted TRA breakpoints."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        diff = int(rng.choice([0, 1, 50, 1000, 1_000_000]))
        scale = max(diff, 3)
        p = int(rng.integers(0, 10_000_000))
        e = p + int(rng.integers(0, 3 * scale))
        jit = lambda: int(rng.choice([0, 0, 1, -1, diff, -diff, diff + 1, -diff - 1, int(rng.integers(-2 * scale, 2 * scale))]))
        sp, se = max(0, p + jit()), max(0, e + jit())
        if kind != "tra" and sp > se:
            sp, se = se, sp                                   # DUP/INV trees hold validated records
        c1, c2 = (str(rng.choice(CHROMS)) for _ in range(2))
        if kind == "tra":
            sc1, sc2 = (c1, c2) if rng.random() < 0.7 else (str(rng.choice(CHROMS)), str(rng.choice(CHROMS)))
            if rng.random() < 0.3:                             # the SV side stored the breakpoints the other way round
                sc1, sc2, sp, se = sc2, sc1, se, sp
            nl = Rec(c1, p, e, "TRA", c2, True, True)
            if rng.random() < 0.3:
                nl = Rec(c2, e, p, "TRA", c1, True, True)
            nl = sv2nl_oracle.validate_record(nl)              # queries are validated, TRA targets are not
            sv = Rec(sc1, sp, se, "BND", sc2, True, True)
        else:
            nl = Rec(c1, p, e, "TDUP" if kind == "dup" else "INV", "", bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
            sv = Rec(c1, sp, se, "DUP" if kind == "dup" else "INV", "", True, True)
        out.append((diff, bool(rng.integers(0, 2)), nl, sv))
    return out


@pytest.mark.parametrize("kind,code,fn", [("dup", sv2nl_ref.DUP, sv2nl_oracle.check_dup),
                                          ("inv", sv2nl_ref.INV, sv2nl_oracle.check_inv),
                                          ("tra", sv2nl_ref.TRA, sv2nl_oracle.check_tra)])
def test_check_condition_equals_the_reference(kind, code, fn):
    seen = set()
    for diff, use_strand, nl, sv in _random_pairs(11 + code, 6000, kind):
        want = sv2nl_ref.check(code, diff, use_strand, nl, sv)
        assert fn(nl, sv, diff, use_strand) == want, (kind, diff, use_strand, nl, sv)
        seen.add(want)
    assert seen == {True, False}


def test_validate_map_key_and_writer_format_equal_the_reference():
    rng = np.random.default_rng(5)
    for _ in range(3000):
        t = str(rng.choice(["TDUP", "INV", "TRA", "BND", "DUP", "INS"]))
        r = Rec(str(rng.choice(CHROMS)), int(rng.integers(0, 2**31)), int(rng.integers(0, 2**31)), t,
                str(rng.choice(CHROMS)) if t in ("TRA", "BND") else "", True, True)
        v = sv2nl_oracle.validate_record(r)
        pos, end, swapped = sv2nl_ref.validate(r)
        assert (v.pos, v.svend) == (pos, end)
        assert swapped == ((v.chrom, v.chr2) != (r.chrom, r.chr2))
        assert sv2nl_oracle.format_map_key(r) == sv2nl_ref.map_key(r)
        assert sv2nl_oracle.format_keys(r) == sv2nl_ref.format_keys(r)


def _cases(tmp_path):
    yield "fixture", NL, SV, 1_000_000, True
    for seed, n_sv, n_nl, diff, strand in ((3, 600, 500, 5000, True), (5, 4000, 3000, 4000, True),
                                           (7, 5000, 4000, 3500, False)):
        d = tmp_path / f"s{seed}"
        d.mkdir()
        nl_path, sv_path = write_synth_vcfs(str(d), seed=seed, n_sv=n_sv, n_nl=n_nl)
        yield f"synthetic{seed}", nl_path, sv_path, diff, strand


def test_whole_mapping_of_the_restatement_equals_the_reference_tool(port_oracle, tmp_path):
    """The reference's run() (three mappers over one pool) vs oracle/sv2nl_oracle.py, as sorted lines: the
    synthetic NL files leave STRAND tags out of some INV records, which exercises the strand carry-over."""
    for name, nl_path, sv_path, diff, strand in _cases(tmp_path):
        want = sv2nl_ref.run(nl_path, sv_path, str(tmp_path / f"ref_{name}"), diff=diff, use_strand=strand)
        got = sv2nl_oracle.sv2nl(port_oracle, read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff=diff,
                                 use_strand=strand)
        for k in ("dup", "inv", "tra"):
            assert sorted(got[k]) == sorted(want[k]), (name, k)
        if name != "fixture":
            assert all(len(want[k]) > 0 for k in want), name


def test_committed_expected_output_is_the_reference_tools_output(tmp_path):
    want = sv2nl_ref.run(NL, SV, str(tmp_path / "fx"))
    assert {k: sorted(v) for k, v in want.items()} == json.load(open(os.path.join(GOLDEN, "sv2nl_expected.json")))


def test_strand_carry_over_of_the_reference_reader(tmp_path):
    """An INV record without STRAND1 keeps both strands of the previous INV record; with STRAND1 but without
    STRAND2 it keeps the previous strand2 (vcf.hpp:305-310 updates one record in place, vcf_info.cpp:17-31
    swallows the exception). The Python reader applies the same carry; the reference shows it through InvMapper."""
    from cases import VCF_INFO_HEADER
    head = ["##fileformat=VCFv4.2", "##contig=<ID=chr1,length=1000000>"] + VCF_INFO_HEADER + \
           ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    sv = head + ["chr1\t1000\ts\tN\t<INV>\t.\t.\tSVTYPE=INV;END=2000"]
    # left-side overlaps (nl.pos <= sv.pos) pass only with strands (+,-)
    nl = head + [
        "chr1\t900\ta\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500;STRAND1=+;STRAND2=-",   # passes
        "chr1\t901\tb\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500",                          # inherits (+,-): passes
        "chr1\t902\tc\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500;STRAND1=-;STRAND2=-",   # fails
        "chr1\t903\td\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500;STRAND2=-",              # STRAND1 missing: inherits (-,-)
        "chr1\t904\te\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500;STRAND1=+",              # (+, inherited -): passes
        "chr1\t905\tf\tN\t<TDUP>\t.\t.\tSVTYPE=TDUP;SVEND=1500",
        "chr1\t906\tg\tN\t<INV>\t.\t.\tSVTYPE=INV;SVEND=1500",                          # still (+,-): passes
    ]
    (tmp_path / "sv.vcf").write_text("\n".join(sv) + "\n")
    (tmp_path / "nl.vcf").write_text("\n".join(nl) + "\n")
    got = sv2nl_ref.run(str(tmp_path / "nl.vcf"), str(tmp_path / "sv.vcf"), str(tmp_path / "o"), diff=10_000)
    assert sorted(l.split("\t")[1] for l in got["inv"]) == ["900", "901", "904", "906"]
    t = read_vcf(str(tmp_path / "nl.vcf"), "nls")
    assert list(t.strand1) == [True, True, False, False, True, True, True]
    assert list(t.strand2) == [False, False, False, False, False, False, False]


# ---- product paths against the reference tool (GPU) ------------------------------------------------------
@pytest.mark.gpu
def test_gpu_python_mapping_equals_the_reference_tool(tmp_path):
    from binary_b200.sv2nl import map_sv2nl
    for name, nl_path, sv_path, diff, strand in _cases(tmp_path):
        want = sv2nl_ref.run(nl_path, sv_path, str(tmp_path / f"ref_{name}"), diff=diff, use_strand=strand)
        got = map_sv2nl(read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff=diff, use_strand=strand)
        for k in ("dup", "inv", "tra"):
            assert sorted(got[k]) == sorted(want[k]), (name, k)


@pytest.mark.gpu
def test_gpu_cpp_tool_equals_the_reference_tool(tmp_path):
    tool_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "standalone", "sv2nl")
    r = subprocess.run(["make", "-C", tool_dir, "sv2nl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for name, nl_path, sv_path, diff, strand in _cases(tmp_path):
        want = sv2nl_ref.run(nl_path, sv_path, str(tmp_path / f"ref_{name}"), diff=diff, use_strand=strand)
        out = str(tmp_path / f"tool_{name}")
        cmd = [os.path.join(tool_dir, "sv2nl"), "--sv", sv_path, "--non-linear", nl_path, "--dis", str(diff), "-o", out]
        if not strand:
            cmd.append("-s")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        for ext in ("dup", "inv", "tra"):
            lines = open(f"{out}.{ext}").read().splitlines()
            assert lines[0] == sv2nl_oracle.HEADER and sorted(lines[1:]) == sorted(want[ext]), (name, ext)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,code", [(1, sv2nl_ref.DUP), (2, sv2nl_ref.INV)])
def test_device_pair_filters_equal_the_reference_predicates(kind, code):
    """bcu_join_filtered (accept<> in join.cu) keeps exactly the overlapping pairs the reference's
    check_condition accepts, pair by pair."""
    from binary_b200 import DeviceIndex
    rng = np.random.default_rng(100 + kind)
    n_t, n_q, diff = 3000, 2000, 700
    tl = rng.integers(0, 300_000, n_t).astype(np.uint32)
    th = (tl + rng.integers(0, 3000, n_t)).astype(np.uint32)
    src = rng.integers(0, n_t, n_q)
    ql = np.maximum(tl[src].astype(np.int64) + rng.integers(-900, 900, n_q), 0).astype(np.uint32)
    qh = np.maximum(th[src].astype(np.int64) + rng.integers(-900, 900, n_q), ql).astype(np.uint32)
    strand = rng.integers(0, 4, n_q).astype(np.uint8)
    ix = DeviceIndex.build(tl, th)
    off, hq, ht = ix.join(ql, qh)
    for use_strand in (True, False):
        goff, ghq, ght = ix.join_filtered(ql, qh, kind=kind, diff=diff, use_strand=use_strand, qstrand=strand)
        got = set(zip(ghq.tolist(), ght.tolist()))
        want = set()
        for q, t in zip(hq.tolist(), ht.tolist()):
            nl = Rec("chr1", int(ql[q]), int(qh[q]), "INV", "", bool(strand[q] & 1), bool(strand[q] & 2))
            sv = Rec("chr1", int(tl[t]), int(th[t]), "INV", "", True, True)
            if sv2nl_ref.check(code, diff, use_strand, nl, sv):
                want.add((q, t))
        assert got == want and 0 < len(want) < hq.size
        assert np.array_equal(np.diff(goff), np.bincount(ghq, minlength=n_q))
    ix.close()
