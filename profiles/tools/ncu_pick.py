#!/usr/bin/env python
"""Print selected metrics of every kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`).
usage: ncu_pick.py report.ncu-rep [substring filter for metric names ...]"""
import csv
import subprocess
import sys

DEFAULT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp", "smsp__warp_issue_stalled", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:90])
        for i, name in enumerate(h):
            if any(p in name for p in pats):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v != 0:
                    print(f"   {name:90s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
