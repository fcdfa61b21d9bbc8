#!/usr/bin/env python
"""Summarises ncu output for profiles/ (run in the CPU container).

  python scripts/ncu_summary.py launches <launches.csv> "<note>"   > profiles/..._launches_summary.txt
  python scripts/ncu_summary.py full <report.ncu-rep> "<note>"     > profiles/..._ncu_full_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    # < 32: instructions issued for a part of the warp (divergence / predication)
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path, note):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print("#", note)
    print("# per-launch times under ncu are cold-cache and serialised: compare SHARES")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} launches={c:4d} total_us={t / 1e3:10.1f} "
              f"avg_us={t / c / 1e3:8.2f} share={100 * t / tot:5.1f}%")
    print(f"total_us={tot / 1e3:.1f}")


def full(path, note):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("#", note)
    for d in data:
        print("-----", d[idx["Kernel Name"]].split("(")[0], "id", d[idx["ID"]])
        for k in KEYS:
            if k in idx:
                print(f"  {k:76s} {d[idx[k]]:>18s} {units[idx[k]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
