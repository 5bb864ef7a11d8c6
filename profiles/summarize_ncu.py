#!/usr/bin/env python
"""Turns an .ncu-rep (brought back in gpurun_out/) into the short JSON summary committed here.
usage: python profiles/summarize_ncu.py <report.ncu-rep> [<launch-index>]"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    d = {"report": rep.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")]}
    stalls = {}
    for i, k in enumerate(hdr):
        if k in KEYS:
            d[k] = "%s %s" % (vals[i], units[i])
        if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio"):
            v = float(vals[i])
            if v >= 0.05:
                stalls[k[len(STALLS):-len("_per_issue_active.ratio")]] = round(v, 2)
    d["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
