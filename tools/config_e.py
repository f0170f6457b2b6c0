"""BASELINE config 5 ("E"): sv2nl end to end on synthetic VCF text -- generate, run the C++ tool
(standalone/sv2nl, parse + GPU join + post-filter + write), and check it against the CPU restatement on a
subsample. usage: python tools/config_e.py [n_sv] [n_nl] [subsample]"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from binary_b200 import synth
from binary_b200.vcf_text import read_vcf

n_sv = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_nl = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
sub = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
names = np.array(synth.HG38_NAMES, dtype=object)
HEAD = ["##fileformat=VCFv4.2"] + [f"##contig=<ID={n},length={l}>" for n, l in synth.HG38] + \
       ["#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]


def write_vcf(path, seed, n, kinds, probs, end_key, nls):
    g, lo, hi = synth.intervals(seed, 0, n, "loguniform", 50, 10_000)
    rng = np.random.default_rng(seed)
    kind = rng.choice(np.array(kinds, dtype=object), size=n, p=probs)
    chr2 = names[rng.integers(0, 24, n)]
    pos2 = rng.integers(1, 40_000_000, n)
    s1 = rng.choice(np.array(["+", "-"], dtype=object), n); s2 = rng.choice(np.array(["+", "-"], dtype=object), n)
    with open(path, "w") as fh:
        fh.write("\n".join(HEAD) + "\n")
        B = 500_000
        for b in range(0, n, B):
            e = min(n, b + B)
            out = []
            for i in range(b, e):
                k, c = kind[i], names[g[i]]
                if k in ("BND", "TRA"):
                    info = f"SVTYPE={k};CHR2={chr2[i]};" + (f"POS2={pos2[i]}" if k == "BND" else f"SVEND={pos2[i]}")
                else:
                    info = f"SVTYPE={k};{end_key}={hi[i] + 1}"
                    if nls:
                        info += f";STRAND1={s1[i]};STRAND2={s2[i]}"
                out.append(f"{c}\t{lo[i] + 1}\tr{i}\tN\t<{k}>\t.\t.\t{info}")
            fh.write("\n".join(out) + "\n")


tmp = tempfile.mkdtemp(prefix="config_e_")
sv_path, nl_path = os.path.join(tmp, "sv.vcf"), os.path.join(tmp, "nl.vcf")
t0 = time.time()
write_vcf(sv_path, 0xE5A0, n_sv, ["DUP", "INV", "BND"], [0.5, 0.25, 0.25], "END", False)
write_vcf(nl_path, 0xE5A1, n_nl, ["TDUP", "INV", "TRA"], [0.5, 0.25, 0.25], "SVEND", True)
print(f"generated {n_sv} SV + {n_nl} NL records in {time.time()-t0:.1f} s "
      f"({os.path.getsize(sv_path)/1e6:.0f} MB + {os.path.getsize(nl_path)/1e6:.0f} MB of VCF text)")
subprocess.run(["make", "-C", os.path.join(ROOT, "standalone", "sv2nl"), "sv2nl"], check=True, capture_output=True)
tool = os.path.join(ROOT, "standalone", "sv2nl", "sv2nl")
out = os.path.join(tmp, "out.tsv")
t0 = time.time()
r = subprocess.run([tool, "--sv", sv_path, "--non-linear", nl_path, "-o", out, "-d"], capture_output=True, text=True)
wall = time.time() - t0
print(r.stderr.strip()); assert r.returncode == 0
print(f"tool wall {wall:.2f} s -> {(n_sv + n_nl)/wall/1e6:.2f} M records/s end to end")
# parity on a subsample: first `sub` NL records vs the whole SV file, through the CPU restatement
import oracle
from oracle import sv2nl_oracle
nl_sub = os.path.join(tmp, "nl_sub.vcf")
with open(nl_path) as src, open(nl_sub, "w") as dst:
    k = 0
    for line in src:
        dst.write(line)
        if not line.startswith("#"):
            k += 1
            if k >= sub: break
sv_sub = os.path.join(tmp, "sv_sub.vcf")
with open(sv_path) as src, open(sv_sub, "w") as dst:
    k = 0
    for line in src:
        dst.write(line)
        if not line.startswith("#"):
            k += 1
            if k >= 4 * sub: break
out2 = os.path.join(tmp, "sub.tsv")
subprocess.run([tool, "--sv", sv_sub, "--non-linear", nl_sub, "-o", out2], check=True)
want = sv2nl_oracle.sv2nl(oracle.Oracle("port"), read_vcf(nl_sub, "nls"), read_vcf(sv_sub, "delly"))
for ext in ("dup", "inv", "tra"):
    got = open(f"{out2}.{ext}").read().splitlines()[1:]
    assert sorted(got) == sorted(want[ext]), ext
    print(f"subsample parity {ext}: {len(got)} lines == oracle")
