"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion uses.

usage: python profiles/summarize.py report.ncu-rep [kernel-substring] > profiles/<name>.txt
       python profiles/summarize.py --traffic report.ncu-rep kind key=value ... [--match substr] [--sum]
The second form appends a record to profiles/traffic.json -- what bench.py reports as `roofline.traffic`
(dram__bytes_read.sum + dram__bytes_write.sum of the matching launch, or with --sum of all matching launches),
keyed by the workload (kind + key=value pairs) and stamped with the report name and the commit."""
import csv
import json
import os
import subprocess
import sys

KEYS = ("gpu__time_duration", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "gpu__dram_throughput",
        "sm__warps_active.avg.pct", "launch__registers", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "lts__t_sector_hit_rate.pct", "shared_mem_per_block_dynamic",
        "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct", "smsp__warps_eligible.avg", "smsp__warps_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed", "sm__inst_executed_pipe_lsu", "smsp__inst_executed_pipe_")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    if sys.argv[1] == "--traffic":
        rep, kind = sys.argv[2], sys.argv[3]
        rest = sys.argv[4:]
        match, do_sum, key = "", False, {}
        i = 0
        while i < len(rest):
            if rest[i] == "--match":
                match = rest[i + 1]; i += 2
            elif rest[i] == "--sum":
                do_sum = True; i += 1
            else:
                k, v = rest[i].split("=", 1)
                key[k] = int(v) if v.lstrip("-").isdigit() else v
                i += 1
        hdr, units, rows = raw_rows(rep)
        ri, wi, ti, ki = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
        sel = [r for r in rows if match in r[ki]]
        if not sel:
            raise SystemExit("no launch matches %r" % match)
        if not do_sum:
            sel = sel[:1]
        total = sum(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in sel)
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        rec = dict(kind=kind, **key, dram_bytes=int(total), launches=len(sel), kernels=sorted({r[ki][:80] for r in sel}),
                   source="%s @ %s" % (os.path.basename(rep), commit))
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
        recs = json.load(open(path)) if os.path.exists(path) else []
        recs = [r for r in recs if not (r.get("kind") == kind and all(r.get(k) == v for k, v in key.items()))] + [rec]
        json.dump(recs, open(path, "w"), indent=1)
        print(json.dumps(rec))
        return
    hdr, units, rows = raw_rows(sys.argv[1])
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for r in rows:
        name = dict(zip(hdr, r)).get("Kernel Name", "?")
        if want not in name:
            continue
        print("== kernel:", name)
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print("  %-90s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
