"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion uses.
usage: python profiles/summarize.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ("gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "gpu__dram_throughput",
        "sm__warps_active.avg.pct", "launch__registers", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate.pct", "shared_mem_per_block_dynamic",
        "sm__pipe_tensor", "sm__throughput.avg.pct", "smsp__warps_eligible.avg", "smsp__warps_active.avg")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = dict(zip(hdr, r)).get("Kernel Name", "?")
    print("== kernel:", name)
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print("  %-90s %-14s %s" % (h, u, v))
