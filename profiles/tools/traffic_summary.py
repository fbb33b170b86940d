#!/usr/bin/env python
"""Per kernel: launches, time, DRAM read+write bytes (ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv).  usage: traffic_summary.py file.csv [launches_per_step]"""
import collections
import csv
import json
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ii, ui = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    d = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
        d.setdefault((int(r[ii]), r[ki].split("(")[0]), {})[r[mi]] = v
    return d


def main():
    d = load(sys.argv[1])
    tot = collections.OrderedDict()
    for (i, k), m in d.items():
        t = tot.setdefault(k, [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += m.get("gpu__time_duration.sum", 0)
        t[2] += m.get("dram__bytes_read.sum", 0)
        t[3] += m.get("dram__bytes_write.sum", 0)
    out = {}
    for k, t in tot.items():
        out[k] = {"launches": t[0], "us": round(t[1], 1), "dram_read_gb": round(t[2] / 1e9, 3), "dram_write_gb": round(t[3] / 1e9, 3),
                  "dram_gbs": round((t[2] + t[3]) / t[1] / 1e3, 1) if t[1] else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
