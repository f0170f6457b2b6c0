"""Turn gpurun_out ncu artefacts into the small, tracked summaries under profiles/.
usage: python tools/summarize_profile.py <tag> <launches.csv> <report.ncu-rep> [workload]"""
import collections, csv, json, os, subprocess, sys

tag, launches_csv, rep = sys.argv[1:4]
workload = sys.argv[4] if len(sys.argv) > 4 else "B"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)

lines = []
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(t for _, t in agg.values())
lines.append(f"# launch list ({os.path.basename(launches_csv)}): ncu --metrics gpu__time_duration.sum --clock-control none")
lines.append("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
lines.append(f"{'launches':>8} {'total_us':>10} {'avg_us':>9} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:8d} {t/1e3:10.1f} {t/n/1e3:9.1f} {100*t/tot:5.1f}%  {k[:110]}")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active"]
lines.append("")
lines.append(f"# ncu --set full --clock-control none ({os.path.basename(rep)}), one block per captured launch")
traffic = {}
for r in rr[2:]:
    for w in want:
        if w in h:
            i = h.index(w)
            lines.append(f"{w:70s} {r[i]} {units[i]}")
    name = r[h.index("Kernel Name")]
    rd = float(r[h.index("dram__bytes_read.sum")]); wr = float(r[h.index("dram__bytes_write.sum")])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    rd *= scale[units[h.index("dram__bytes_read.sum")]]; wr *= scale[units[h.index("dram__bytes_write.sum")]]
    traffic[name] = rd + wr
    lines.append(f"{'dram traffic (read+write) bytes':70s} {rd + wr:.0f}")
    lines.append("--")
open(os.path.join(out_dir, f"{tag}_summary.txt"), "w").write("\n".join(lines) + "\n")
# bench.py reads this: dram bytes per join step = probe + emit launches
tj = os.path.join(out_dir, "traffic.json")
cur = json.load(open(tj)) if os.path.exists(tj) else {}
cur[workload] = {"bytes_per_step": sum(traffic.values()), "per_kernel": traffic, "source": f"profiles/{tag}_summary.txt"}
json.dump(cur, open(tj, "w"), indent=1)
print("\n".join(lines))
