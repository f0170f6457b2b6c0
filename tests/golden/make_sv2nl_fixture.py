"""Builds tests/golden/nl_fixture.vcf from the reference's NL fixture (test/data/debug.vcf.gz: 2 TRA,
1 INS, 1 TDUP chr14, 2 TDUP chr17 with POS > SVEND) keeping only what sv2nl reads: the contig header
lines (the 25 primary contigs plus three '_' contigs to exercise the filter of mapper.hpp:241-243), and per
record CHROM, POS, ID and INFO SVTYPE/CHR2/SVEND/STRAND1/STRAND2. Run in the authoring container only.
The SV side (tests/golden/sv_fixture.vcf) is authored by hand: the reference ships no delly VCF."""
import gzip
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/test/data/debug.vcf.gz"
KEEP = ("SVTYPE", "CHR2", "SVEND", "STRAND1", "STRAND2")
INFO_HEADER = ['##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">', '##INFO=<ID=CHR2,Number=1,Type=String,Description="Chromosome for the second breakpoint">', '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the structural variant">', '##INFO=<ID=POS2,Number=1,Type=Integer,Description="Position of the second breakpoint (BND)">', '##INFO=<ID=SVEND,Number=1,Type=Integer,Description="2nd position of the structural variant">', '##INFO=<ID=STRAND1,Number=1,Type=String,Description="Strand for breakpoint1">', '##INFO=<ID=STRAND2,Number=1,Type=String,Description="Strand for breakpoint2">']

out = ["##fileformat=VCFv4.3", "##source=ScanNLS (reduced copy of BINARY test/data/debug.vcf.gz)"]
extra = 0
with gzip.open(SRC, "rt") as fh:
    for line in fh:
        if line.startswith("##contig"):
            name = re.search(r"ID=([^,>]+)", line).group(1)
            if "_" not in name:
                out.append(line.strip())
            elif extra < 3:
                out.append(line.strip()); extra += 1
        elif line.startswith("#CHROM"):
            out.extend(INFO_HEADER)   # htslib types INFO values by these lines (undeclared tags become String)
            out.append("\t".join(line.strip().split("\t")[:8]))
        elif not line.startswith("#"):
            c = line.rstrip("\n").split("\t")
            info = dict(kv.split("=", 1) for kv in c[7].split(";") if "=" in kv)
            c[7] = ";".join(f"{k}={info[k]}" for k in KEEP if k in info)
            out.append("\t".join(c[:8]))
open(os.path.join(HERE, "nl_fixture.vcf"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if not l.startswith("##contig")))
