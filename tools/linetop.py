#!/usr/bin/env python
"""Top source lines of one kernel by executed warp instructions (tools/lineagg.py output joined with the source).
usage: linetop.py <sass from nvdisasm -g> <ncu --page source --csv> <kernel tag> <source file> [n]"""
import re, subprocess, sys
sass, csvf, tag, src = sys.argv[1:5]
n = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run([sys.executable, __file__.replace("linetop.py", "lineagg.py"), sass, csvf, tag, src.split("/")[-1]],
                     capture_output=True, text=True).stdout
rows = []
for l in out.split("\n"):
    m = re.match(r"\('([^']+)', (\d+)\)\s+samples\s+(\d+) \(\s*([\d.]+)%\)\s+winst\s+(\d+) \(\s*([\d.]+)%\)", l)
    if m:
        rows.append((float(m.group(6)), float(m.group(4)), m.group(1), int(m.group(2))))
    elif l.startswith("totals"):
        print(l)
text = open(src).read().split("\n")
for inst, samp, f, ln in sorted(rows, reverse=True)[:n]:
    t = text[ln - 1].strip()[:105] if f == src.split("/")[-1] and ln <= len(text) else f
    print("inst %5.1f%% samp %5.1f%%  %s:%d  %s" % (inst, samp, f, ln, t))
