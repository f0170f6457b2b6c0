"""GPU (-m gpu): BASELINE.json configs[4] ("E") -- sv2nl end to end on synthetic VCF text. The C++ tool
(standalone/sv2nl: parse once, batched GPU joins with fused DUP / INV filters, duplicate-key rule, writer) against
the UNMODIFIED reference tool (oracle/_ref/libsv2nl_ref.so, see tests/test_sv2nl_reference.py) on the FULL set, by
line count and an order-independent hash of the lines of each of the three output files (the reference does not
define the order: its tasks interleave).

Default size 60 k SV x 300 k NL records so that the suite stays within minutes (the reference tool re-parses both
files in every chromosome task, mapper.hpp:196-197). BCU_TEST_FULL_E=1 runs the configuration's own 1 M x 5 M for the
DUP and INV files (profiles/r02_config_e.txt holds that run); the reference's TraMapper is left out at that size: it
joins on the raw [POS, POS2] intervals of ALL BND records (mapper.cpp:103), so an NL record visits a large fraction
of them -- 1.25 M x 250 k = 3e11 pair visits, hours on the host (the first attempt at it was stopped after 35 min)."""
import os
import subprocess
import time

import pytest

from oracle import sv2nl_ref
from cases import line_set_digest, write_config_e_vcfs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not sv2nl_ref.available(), reason="reference sv2nl library not built")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_e_full_set_count_and_hash(tmp_path):
    full = bool(os.environ.get("BCU_TEST_FULL_E"))
    n_sv, n_nl = (1_000_000, 5_000_000) if full else (60_000, 300_000)
    nl_path, sv_path = write_config_e_vcfs(str(tmp_path), n_sv, n_nl)
    tool_dir = os.path.join(ROOT, "standalone", "sv2nl")
    r = subprocess.run(["make", "-C", tool_dir, "sv2nl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = str(tmp_path / "tool")
    t0 = time.time()
    r = subprocess.run([os.path.join(tool_dir, "sv2nl"), "--sv", sv_path, "--non-linear", nl_path, "-o", out],
                       capture_output=True, text=True, timeout=1200)
    t_tool = time.time() - t0
    assert r.returncode == 0, r.stderr
    t0 = time.time()
    exts = ("dup", "inv") if full else ("dup", "inv", "tra")
    want = sv2nl_ref.run(nl_path, sv_path, str(tmp_path / "ref"), threads=os.cpu_count() or 4, mappers=exts)
    t_ref = time.time() - t0
    report = [f"config E {n_sv} SV x {n_nl} NL records: tool (all three mappers) {t_tool:.2f} s, reference tool "
              f"({'+'.join(exts)}) {t_ref:.2f} s ({os.cpu_count()} host threads)"]
    for ext in ("dup", "inv", "tra"):
        got = open(f"{out}.{ext}").read().splitlines()
        assert got[0] == "chrom\tpos\tend\tsvtype\tchrom\tpos\tend\tsvtype"
        if ext not in exts:
            report.append(f"  .{ext}: {len(got) - 1} lines (not compared at this size, see the module docstring)")
            continue
        g, w = line_set_digest(got[1:]), line_set_digest(want[ext])
        report.append(f"  .{ext}: {g[0]} lines, hash {g[1]:016x} (reference: {w[0]} lines, {w[1]:016x})")
        assert g == w, ext
    assert sum(len(want[e]) for e in exts) > 0
    print("\n".join(report))
    if os.environ.get("BCU_REPORT_DIR"):
        with open(os.path.join(os.environ["BCU_REPORT_DIR"], "config_e.txt"), "w") as fh:
            fh.write("\n".join(report) + "\n")
