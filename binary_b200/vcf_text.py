"""Minimal text-VCF reader for the sv2nl path (plain or gzip/bgzip), columnar output.

The reference parses through htslib (``library/include/binary/parser/vcf.hpp``), which is not in this
image; the north star keeps that parser on the host, so this is only the shim SURVEY.md section 8c asks
for at that boundary. It honours exactly the fields sv2nl reads:

* ``chrom``  = column 1; ``pos`` = POS - 1 (0-based, ``vcf.hpp:305-310``, ``test_vcf.cpp:100``)
* INFO ``SVTYPE`` (required), ``CHR2`` for TRA/BND, ``STRAND1``/``STRAND2`` == "+" for INV. The reference
  updates ONE record object in place while it iterates (``vcf.hpp:262-275,305-310``) and swallows the
  exception of a missing strand tag (``vcf_info.cpp:17-31``), so an INV record without ``STRAND1`` keeps BOTH
  strands of the previous INV record of the file (initially "+","+"), and one with ``STRAND1`` but without
  ``STRAND2`` keeps the previous ``strand2``. The same carry is applied here (pinned against the unmodified
  reference in tests/test_sv2nl_reference.py). End coordinate = ``POS2`` if SVTYPE == BND, else ``SVEND`` when the source is "nls", else
  ``END`` (``standalone/sv2nl/source/vcf_info.cpp:9-43``); the end stays the raw 1-based INFO integer while
  ``pos`` is 0-based -- the reference does not reconcile them and neither do we.
* contig list = ``##contig=<ID=...>`` header lines in order (``vcf.hpp:577-589``)

Parse-once: both files are read ONE time into SoA arrays (the reference re-parses both files in every
chromosome task, ``mapper.hpp:196-197``).
"""
from __future__ import annotations

import gzip
import re
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

SVTYPES = ("DUP", "TDUP", "INV", "TRA", "BND", "INS", "DEL", "IDUP")


class VcfReaderError(ValueError):
    """Mirror of ``binary::VcfReaderError`` (``library/include/binary/exception.hpp:13``)."""


@dataclass
class VcfTable:
    contigs: List[str]            # header order
    chrom: np.ndarray             # object array of chromosome names, one per record
    pos: np.ndarray               # u32, 0-based
    svend: np.ndarray             # u32, raw INFO integer
    svtype: np.ndarray            # object array of SVTYPE strings
    chr2: np.ndarray              # object array ("" unless TRA/BND)
    strand1: np.ndarray           # bool, True = "+"
    strand2: np.ndarray

    def __len__(self) -> int:
        return int(self.pos.size)


_INFO_RE = re.compile(r"(?:^|;)([^=;]+)=([^;]*)")


def _open(path: str):
    with open(path, "rb") as fh:
        magic = fh.read(2)
    return gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")


def read_vcf(path: str, source: str) -> VcfTable:
    """``source`` = "nls" (ScanNLS non-linear calls, end in SVEND) or "delly" (end in END / POS2)."""
    contigs: List[str] = []
    chrom, pos, svend, svtype, chr2, s1, s2 = [], [], [], [], [], [], []
    carry1 = carry2 = True  # Sv2nlInfoField's defaults (vcf_info.hpp:20-21)
    with _open(path) as fh:
        for line_no, line in enumerate(fh, 1):
            if line.startswith("##contig=<"):
                m = re.search(r"ID=([^,>]+)", line)
                if m:
                    contigs.append(m.group(1))
                continue
            if line.startswith("#") or not line.strip():
                continue
            cols = line.rstrip("\n").split("\t")
            if len(cols) < 8:
                raise VcfReaderError(f"{path}:{line_no}: fewer than 8 columns")
            info: Dict[str, str] = dict(_INFO_RE.findall(cols[7]))
            if "SVTYPE" not in info:
                raise VcfReaderError(f"{path}:{line_no}: INFO/SVTYPE missing")
            t = info["SVTYPE"]
            end_key = "POS2" if t == "BND" else ("SVEND" if source == "nls" else "END")
            if end_key not in info:
                raise VcfReaderError(f"{path}:{line_no}: INFO/{end_key} missing")
            c2 = ""
            if t in ("TRA", "BND"):
                if "CHR2" not in info:
                    raise VcfReaderError(f"{path}:{line_no}: INFO/CHR2 missing")
                c2 = info["CHR2"]
            if t == "INV" and "STRAND1" in info:  # a missing tag keeps what the previous INV record left
                carry1 = info["STRAND1"] == "+"
                if "STRAND2" in info:
                    carry2 = info["STRAND2"] == "+"
            st1, st2 = carry1, carry2
            chrom.append(cols[0])
            pos.append(int(cols[1]) - 1)
            svend.append(int(info[end_key]))
            svtype.append(t)
            chr2.append(c2)
            s1.append(st1)
            s2.append(st2)
    u32 = lambda a: np.array(a, dtype=np.int64).astype(np.uint32)
    obj = lambda a: np.array(a, dtype=object)
    return VcfTable(contigs, obj(chrom), u32(pos), u32(svend), obj(svtype), obj(chr2),
                    np.array(s1, dtype=bool), np.array(s2, dtype=bool))
