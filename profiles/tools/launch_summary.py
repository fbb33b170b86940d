#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel
count, total and mean duration, share of the total.  usage: launch_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        d[r[ki].split("(")[0][:60]].append(v * scale)
    tot = sum(sum(v) for v in d.values())
    print(f"{path}: {sum(len(v) for v in d.values())} launches, {tot / 1e3:.2f} ms in kernels")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {k:60s} n={len(v):5d} total={sum(v):10.1f} us  mean={sum(v) / len(v):8.1f} us  min={min(v):8.1f}  share={100 * sum(v) / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
