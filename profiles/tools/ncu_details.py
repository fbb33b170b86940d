#!/usr/bin/env python
"""Pick the headline metrics out of an `ncu -i x.ncu-rep --page details --csv` export, one line per launch.
usage: ncu_details.py file.csv [kernel substring]"""
import collections
import csv
import sys

WANT = ["Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "L2 Hit Rate", "L1/TEX Hit Rate",
        "Achieved Occupancy", "Theoretical Occupancy", "Registers Per Thread", "Executed Ipc Active",
        "Dynamic Shared Memory Per Block", "Block Limit Shared Mem", "Block Limit Registers", "Mem Busy", "Max Bandwidth"]


def main(path, sub=""):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    ix = {n: i for i, n in enumerate(h)}
    d = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        key = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0], r[ix["Grid Size"]], r[ix["Block Size"]])
        d.setdefault(key, {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]] + " " + r[ix["Metric Unit"]]
    for k, m in d.items():
        if sub in k[1]:
            print(k[0], k[1], k[2], k[3])
            print("    " + "; ".join(f"{w}: {m[w].strip()}" for w in WANT if w in m))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
