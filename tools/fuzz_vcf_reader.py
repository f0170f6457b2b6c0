"""ASAN/UBSAN run of the sv2nl tool's VCF reader (standalone/sv2nl/vcf_text.hpp) over adversarial inputs:
9 MB lines, 200 k INFO keys, missing newline at the end, CRLF, NUL bytes, empty values, overflowing numbers,
> 256 SVTYPE values, random byte mutations -- each with 1 and 4 parser threads and both sources. The reader may
accept (exit 0) or reject with a VcfReaderError (exit 1); anything else, or a sanitizer report, fails.
CPU only. usage: python tools/fuzz_vcf_reader.py"""
import os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp(prefix="fuzz_vcf_")
exe = os.path.join(tmp, "dump_asan")
subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                os.path.join(ROOT, "tests", "cpp", "dump_vcf.cpp"), "-lz", "-lpthread"], check=True)
random.seed(3)
head = "##contig=<ID=chr1,length=1000>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
good = "chr1\t5\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n"
cases = {
    "long_line": head + "chr1\t5\t" + "x" * (9 << 20) + "\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n" + good * 10,
    "long_info": head + "chr1\t5\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;" + ";".join(f"K{i}=v" for i in range(200000)) + ";END=9\n" + good,
    "no_newline_end": head + good * 3 + good.rstrip("\n"),
    "only_header": head,
    "empty_info_values": head + "chr1\t5\t.\tN\t<DUP>\t.\t.\tSVTYPE=;END=\n",
    "tabs_only": head + "\t\t\t\t\t\t\t\n",
    "eq_only": head + "chr1\t5\t.\tN\t<DUP>\t.\t.\t=;==;;;SVTYPE=DUP;END=9;\n",
    "huge_numbers": head + "chr1\t99999999999999999999\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n",
    "negative": head + "chr1\t0\t.\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=-5\n",
    "crlf": (head + good * 5).replace("\n", "\r\n"),
    "nul_bytes": head + "chr1\t5\t.\tN\t<D\0P>\t.\t.\tSVTYPE=D\0P;END=9\n" + good,
    "many_types": head + "".join(f"chr1\t5\t.\tN\t<T>\t.\t.\tSVTYPE=T{i};END=9\n" for i in range(300)),
}
for i in range(30):
    b = bytearray((head + good * 200).encode())
    for _ in range(20):
        b[random.randrange(len(b))] = random.randrange(256)
    cases[f"mut{i}"] = bytes(b)
bad = 0
for name, data in cases.items():
    p = os.path.join(tmp, f"{name}.vcf")
    with open(p, "wb") as fh:
        fh.write(data if isinstance(data, bytes) else data.encode())
    for thr in ("1", "4"):
        for src in ("nls", "delly"):
            r = subprocess.run([exe, p, src], capture_output=True, env=dict(os.environ, SV2NL_PARSE_THREADS=thr))
            if r.returncode not in (0, 1) or b"ERROR" in r.stderr or b"runtime error" in r.stderr:
                bad += 1
                print(name, thr, src, r.returncode, r.stderr[-400:])
    os.remove(p)
print(f"sanitizer fuzz: {len(cases)} inputs x 2 thread counts x 2 sources, failures: {bad}")
sys.exit(1 if bad else 0)
