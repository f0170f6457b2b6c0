"""TEST INFRASTRUCTURE ONLY -- ctypes front end of ``oracle/_ref/libsv2nl_ref.so``: the UNMODIFIED
reference sv2nl sources (``standalone/sv2nl/source/{mapper,vcf_info,writer}.cpp``, ``include/*.hpp``,
``library/include/binary/parser/vcf.hpp``) compiled in the authoring container over the text-VCF stand-in
for htslib (``oracle/stubs/``), see ``oracle/sv2nl_ref_harness.cpp``. It is what pins the sv2nl level:

* :func:`check` -- ``{Dup,Inv,Tra}Mapper::check_condition`` (mapper.cpp:50-79,144-156)
* :func:`validate` -- ``validate_record`` (helper.hpp:52-63)
* :func:`map_key` / :func:`format_keys` -- ``format_map_key`` (helper.hpp:84-91), ``Writer::format_keys``
* :func:`run` -- the tool's ``run()`` (main.cpp:47-84): the three mappers over one pool, three TSV files

The prebuilt library travels to the GPU box (git-ignored, not gpurun-ignored); nothing here reads
/root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libsv2nl_ref.so")
DUP, INV, TRA = 0, 1, 2
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(SO)
        s, u, i = C.c_char_p, C.c_uint32, C.c_int
        lib.sv2nl_ref_check.restype = i
        lib.sv2nl_ref_check.argtypes = [i, u, i, s, u, u, s, s, i, i, s, u, u, s, s]
        lib.sv2nl_ref_validate.restype = i
        lib.sv2nl_ref_validate.argtypes = [s, u, u, s, s, C.POINTER(u), C.POINTER(u), C.POINTER(i)]
        for name in ("sv2nl_ref_map_key", "sv2nl_ref_format_keys"):
            getattr(lib, name).restype = i
            getattr(lib, name).argtypes = [s, u, u, s, s, C.c_char_p, i]
        lib.sv2nl_ref_run.restype = i
        lib.sv2nl_ref_run.argtypes = [s, s, s, u, i, i]
        lib.sv2nl_ref_run_some.restype = i
        lib.sv2nl_ref_run_some.argtypes = [s, s, s, u, i, i, i]
        _lib = lib
    return _lib


def _b(x: str) -> bytes:
    return x.encode()


def check(kind: int, diff: int, use_strand: bool, nl, sv) -> bool:
    """nl / sv: objects with chrom, pos, svend, svtype, chr2 (+ strand1, strand2 on nl)."""
    r = _load().sv2nl_ref_check(kind, diff, int(use_strand), _b(nl.chrom), nl.pos, nl.svend, _b(nl.svtype),
                                _b(nl.chr2), int(nl.strand1), int(nl.strand2), _b(sv.chrom), sv.pos, sv.svend,
                                _b(sv.svtype), _b(sv.chr2))
    assert r in (0, 1)
    return bool(r)


def validate(r):
    """(pos, svend, chroms_swapped) after validate_record."""
    p, e, sw = C.c_uint32(), C.c_uint32(), C.c_int()
    _load().sv2nl_ref_validate(_b(r.chrom), r.pos, r.svend, _b(r.svtype), _b(r.chr2), C.byref(p), C.byref(e),
                               C.byref(sw))
    return p.value, e.value, bool(sw.value)


def _text(fn, r) -> str:
    buf = C.create_string_buffer(512)
    n = fn(_b(r.chrom), r.pos, r.svend, _b(r.svtype), _b(r.chr2), buf, 512)
    assert n < 512
    return buf.value.decode()


def map_key(r) -> str:
    return _text(_load().sv2nl_ref_map_key, r)


def format_keys(r) -> str:
    return _text(_load().sv2nl_ref_format_keys, r)


def run(nl_path: str, sv_path: str, out_prefix: str, diff: int = 1_000_000, threads: int = 4,
        use_strand: bool = True, mappers=("dup", "inv", "tra")) -> Dict[str, List[str]]:
    """Runs the reference's mappers (all three by default); returns the DATA lines of <prefix>.dup/.inv/.tra (header
    checked; a mapper that was left out has none)."""
    mask = sum(bit for name, bit in (("dup", 1), ("inv", 2), ("tra", 4)) if name in mappers)
    rc = _load().sv2nl_ref_run_some(_b(nl_path), _b(sv_path), _b(out_prefix), diff, threads, int(use_strand), mask)
    if rc != 0:
        raise RuntimeError("reference sv2nl run failed")
    out = {}
    for ext in ("dup", "inv", "tra"):
        lines = open(f"{out_prefix}.{ext}").read().splitlines()
        assert lines and lines[0] == "chrom\tpos\tend\tsvtype\tchrom\tpos\tend\tsvtype"
        out[ext] = lines[1:]
    return out
