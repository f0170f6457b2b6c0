"""sv2nl level (BASELINE config 1 + SURVEY 8f): text-VCF shim, the CPU restatement of the mapping loop
(oracle/sv2nl_oracle.py, parity UNPINNED at this level -- the reference has no sv2nl tests), and the
batched GPU mapping (binary_b200/sv2nl.py) against it."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import sv2nl_oracle
from binary_b200.vcf_text import VcfReaderError, read_vcf
from cases import write_synth_vcfs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NL, SV = os.path.join(GOLDEN, "nl_fixture.vcf"), os.path.join(GOLDEN, "sv_fixture.vcf")


def test_reader_on_the_reference_nl_fixture():
    # what the reference's parser test pins (test/source/test_parser/test_vcf.cpp:88-101,158-174)
    nl = read_vcf(NL, "nls")
    assert len(nl) == 6                                   # tree of all records has size 6
    assert nl.chrom[0] == "chr10" and nl.svtype[0] == "TRA" and nl.pos[0] == 93567288 - 1
    assert int((nl.svtype == "TRA").sum()) == 2           # TRA-filtered tree has size 2
    assert nl.chr2[0] == "chr17" and nl.svend[0] == 7705262
    assert [c for c in nl.contigs if "_" not in c][:3] == ["chr1", "chr10", "chr11"] and len(nl.contigs) == 28
    assert nl.svend[4] < nl.pos[4]                        # chr17 TDUP with POS > SVEND (validate_record case)
    sv = read_vcf(SV, "delly")
    assert len(sv) == 16 and sv.svend[10] == 7705300 and sv.chr2[10] == "chr17"   # BND: POS2 / CHR2


def test_reader_errors(tmp_path):
    p = tmp_path / "bad.vcf"
    p.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\nchr1\t5\t.\tN\t<DUP>\t.\t.\tEND=9\n")
    with pytest.raises(VcfReaderError):
        read_vcf(str(p), "delly")
    p.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\nchr1\t5\t.\tN\t<BND>\t.\t.\tSVTYPE=BND;CHR2=chr2\n")
    with pytest.raises(VcfReaderError):
        read_vcf(str(p), "delly")


def test_oracle_restatement_on_fixture_matches_committed_output(port_oracle):
    res = sv2nl_oracle.sv2nl(port_oracle, read_vcf(NL, "nls"), read_vcf(SV, "delly"))
    want = json.load(open(os.path.join(GOLDEN, "sv2nl_expected.json")))
    assert {k: sorted(v) for k, v in res.items()} == want
    assert len(want["dup"]) == 5 and len(want["tra"]) == 4 and want["inv"] == []
    # duplicate NL key (the two chr17 TDUP records) is written once: 3 SV hits, not 6
    assert sum(l.startswith("chr17\t7708250") for l in want["dup"]) == 3


def test_oracle_restatement_on_synthetic_inputs(port_oracle, tmp_path):
    nl_path, sv_path = write_synth_vcfs(str(tmp_path), seed=3, n_sv=600, n_nl=500)
    res = sv2nl_oracle.sv2nl(port_oracle, read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff=5000)
    assert all(len(res[k]) > 0 for k in ("dup", "inv", "tra"))
    # NL records on '_' contigs are never mapped (mapper.hpp:241-243); the first column is the NL chrom
    assert not any("_" in l.split("\t")[0].split(",")[0] for k in res for l in res[k])
    loose = sv2nl_oracle.sv2nl(port_oracle, read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff=5000,
                               use_strand=False)
    assert len(loose["inv"]) >= len(res["inv"]) and sorted(loose["dup"]) == sorted(res["dup"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["fixture", "synthetic", "synthetic-nostrand"])
def test_gpu_mapping_equals_oracle(port_oracle, tmp_path, case):
    from binary_b200.sv2nl import HEADER, map_sv2nl, run
    if case == "fixture":
        nl_path, sv_path, diff, strand = NL, SV, 1_000_000, True
    else:
        nl_path, sv_path = write_synth_vcfs(str(tmp_path), seed=5, n_sv=4000, n_nl=3000)
        diff, strand = 4000, case == "synthetic"
    nl, sv = read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly")
    want = sv2nl_oracle.sv2nl(port_oracle, nl, sv, diff=diff, use_strand=strand)
    got = map_sv2nl(nl, sv, diff=diff, use_strand=strand)
    for k in ("dup", "inv", "tra"):
        assert sorted(got[k]) == sorted(want[k]), k
    counts = run(nl_path, sv_path, str(tmp_path / "out.tsv"), diff=diff, use_strand=strand)
    assert counts == {k: len(v) for k, v in want.items()}
    assert open(tmp_path / "out.tsv.dup").readline().rstrip("\n") == HEADER == sv2nl_oracle.HEADER


# ---- the C++ tool's reader (standalone/sv2nl/vcf_text.hpp) on CPU: same table as the Python reader ------
CPP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")


@pytest.fixture(scope="module")
def dump_vcf():
    r = subprocess.run(["make", "-C", CPP_DIR, "dump_vcf"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(CPP_DIR, "dump_vcf")


def _python_dump(path, source):
    t = read_vcf(path, source)
    lines = [f"contig\t{c}" for c in t.contigs]
    for i in range(len(t)):
        lines.append(f"rec\t{t.chrom[i]}\t{int(t.pos[i])}\t{int(t.svend[i])}\t{t.svtype[i]}\t{t.chr2[i]}\t"
                     f"{int(t.strand1[i])}\t{int(t.strand2[i])}")
    return lines


@pytest.mark.parametrize("threads", ["1", "5"])
def test_cpp_reader_equals_python_reader(dump_vcf, tmp_path, threads):
    nl_path, sv_path = write_synth_vcfs(str(tmp_path), seed=5, n_sv=4000, n_nl=6000)
    gz = str(tmp_path / "nl_copy.vcf.gz")
    import gzip
    with open(nl_path, "rb") as src, gzip.open(gz, "wb") as dst:
        dst.write(src.read())
    env = dict(os.environ, SV2NL_PARSE_THREADS=threads)
    for path, source in ((NL, "nls"), (SV, "delly"), (nl_path, "nls"), (sv_path, "delly"), (gz, "nls")):
        r = subprocess.run([dump_vcf, path, source], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines() == _python_dump(path, source), path


def test_cpp_reader_reports_the_offending_line(dump_vcf, tmp_path):
    head = "##contig=<ID=chr1,length=1000>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    good = "chr1\t5\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n"
    cases = {
        "INFO/SVTYPE missing": "chr1\t5\t.\tN\t<DUP>\t.\t.\tEND=9\n",
        "INFO/POS2 missing": "chr1\t5\t.\tN\t<BND>\t.\t.\tSVTYPE=BND;CHR2=chr2\n",
        "INFO/CHR2 missing": "chr1\t5\t.\tN\t<BND>\t.\t.\tSVTYPE=BND;POS2=7\n",
        "fewer than 8 columns": "chr1\t5\t.\tN\n",
        "POS is not a number": "chr1\tfive\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n",
    }
    for what, bad in cases.items():
        p = tmp_path / "bad.vcf"
        p.write_text(head + good * 700 + bad + good * 3)       # line 703 of the file, deep inside a block
        for threads in ("1", "4"):
            r = subprocess.run([dump_vcf, str(p), "delly"], capture_output=True, text=True,
                               env=dict(os.environ, SV2NL_PARSE_THREADS=threads))
            assert r.returncode == 1 and f"bad.vcf:703: {what}" in r.stderr, (what, threads, r.stderr)
    # a last line without newline, CRLF line ends and an empty file are all fine
    p = tmp_path / "tail.vcf"
    p.write_text(head + good.replace("\n", "\r\n") + good.rstrip("\n"))
    r = subprocess.run([dump_vcf, str(p), "delly"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.count("rec\t") == 2 and "\r" not in r.stdout
    p.write_text("")
    assert subprocess.run([dump_vcf, str(p), "delly"], capture_output=True, text=True).stdout == ""


# ---- the C++ tool (standalone/sv2nl): same CLI as the reference's sv2nl ---------------------------------
import subprocess

TOOL_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "standalone", "sv2nl")


@pytest.fixture(scope="module")
def sv2nl_tool():
    r = subprocess.run(["make", "-C", TOOL_DIR, "sv2nl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return os.path.join(TOOL_DIR, "sv2nl")


def test_cpp_tool_builds_and_prints_help(sv2nl_tool):
    # the reference CI's only sv2nl check is `sv2nl -h` (.github/workflows/linux.yml:109-111)
    r = subprocess.run([sv2nl_tool, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "--non-linear" in r.stdout and "--sv" in r.stdout
    assert subprocess.run([sv2nl_tool], capture_output=True).returncode == 2
    bad = subprocess.run([sv2nl_tool, "--sv", "/nonexistent.vcf", "--non-linear", NL], capture_output=True, text=True)
    assert bad.returncode == 1 and "cannot open" in bad.stderr      # fails loudly, unlike the reference's pool


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["fixture", "synthetic", "synthetic-nostrand"])
def test_cpp_tool_output_equals_oracle(sv2nl_tool, port_oracle, tmp_path, case):
    if case == "fixture":
        nl_path, sv_path, diff, strand = NL, SV, 1_000_000, True
    else:
        nl_path, sv_path = write_synth_vcfs(str(tmp_path), seed=7, n_sv=5000, n_nl=4000)
        diff, strand = 3500, case == "synthetic"
    want = sv2nl_oracle.sv2nl(port_oracle, read_vcf(nl_path, "nls"), read_vcf(sv_path, "delly"), diff=diff,
                              use_strand=strand)
    out = str(tmp_path / "out.tsv")
    cmd = [sv2nl_tool, "--sv", sv_path, "--non-linear", nl_path, "--dis", str(diff), "-o", out, "-t", "4"]
    if not strand:
        cmd.append("-s")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for ext in ("dup", "inv", "tra"):
        lines = open(f"{out}.{ext}").read().splitlines()
        assert lines[0] == sv2nl_oracle.HEADER
        assert sorted(lines[1:]) == sorted(want[ext]), ext
