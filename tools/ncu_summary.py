#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): per captured launch the few counters that decide where the time
goes -- duration, DRAM traffic, pipe/issue utilisation, occupancy, shared-memory conflicts, and the top warp
stall reasons. Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x_summary.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("=" * 100)
        print(r[name_i][:90], " grid", r[hdr.index("Grid Size")], " block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]}")
        stalls = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and r[i]]
        # fall back to the pcsamp-independent counters
        if not stalls:
            stalls = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled") and r[i]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''):70s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
