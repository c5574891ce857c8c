#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` output into the handful of figures DESIGN.md and bench.py quote:
per kernel (averaged over its captured launches) duration, DRAM bytes, L2 hit rate, achieved occupancy,
issue-slot utilisation, warp execution efficiency and the top stall reasons.
    python pacbio_b200/tools/ncu_summary.py raw.csv [kernel-substring ...] > summary.json"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_sector_hit_rate.pct": "l2_sector_hit_rate_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "l1tex__t_sector_hit_rate.pct": "l1_sector_hit_rate_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilization_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed": "memory_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
}
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    path, filters = sys.argv[1], sys.argv[2:]
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    units = rows[1]                                              # second line: the unit of every column
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "second": 1e9, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "s": 1e9, "ms": 1e6, "us": 1e3, "ns": 1.0,
             "Gbyte/second": 1e9, "Mbyte/second": 1e6, "Tbyte/second": 1e12}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    ik = hdr.index("Kernel Name")
    out = {}
    for r in body:
        name = re.sub(r"\(.*", "", r[ik]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        if filters and not any(f in name for f in filters):
            continue
        k = out.setdefault(name, {"launches": 0, "sum": {}, "stalls": {}})
        k["launches"] += 1
        for i, h in enumerate(hdr):
            v = num(r[i])
            if v is None:
                continue
            if h in WANT:
                v *= scale.get(units[i], 1.0)                    # bytes and nanoseconds
                k["sum"][WANT[h]] = k["sum"].get(WANT[h], 0.0) + v
            m = STALL.match(h)
            if m:
                s = m.group(1) or m.group(2)
                k["stalls"][s] = k["stalls"].get(s, 0.0) + v
    res = {}
    for name, k in out.items():
        n = k["launches"]
        d = {a: b / n for a, b in k["sum"].items()}
        d["launches_captured"] = n
        if "dram_read_bytes" in d:
            d["dram_bytes_per_launch"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
            if d.get("duration_ns"):
                d["dram_gbs"] = d["dram_bytes_per_launch"] / d["duration_ns"]
        if "threads_per_instruction" in d:
            d["warp_execution_efficiency_pct"] = 100.0 * d["threads_per_instruction"] / 32.0
        top = sorted(k["stalls"].items(), key=lambda x: -x[1])[:4]
        d["top_stalls_warps_per_issue"] = {a: round(b / n, 3) for a, b in top}
        res[name] = d
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
