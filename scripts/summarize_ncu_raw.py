#!/usr/bin/env python
"""Summarise an `ncu --set full ... --page raw --csv` export: one line per kernel launch with duration, DRAM bytes,
DRAM / SM / tensor-pipe utilisation, occupancy and the top warp-stall reasons."""
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    units = next(rd)
    ki = hdr.index("Kernel Name")
    scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9, "s": 1e9,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def col(name):
        return hdr.index(name) if name in hdr else None
    cols = {"dur_ns": col("gpu__time_duration.sum"), "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"),
            "dram%": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "sm%": col("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor%": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "tc%": col("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue%": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "occ%": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "regs": col("launch__registers_per_thread")}
    stalls = [(i, h) for i, h in enumerate(hdr) if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$", h)]

    def f(r, k):
        i = cols[k]
        try:
            return float(r[i].replace(",", "")) * scale.get(units[i], 1.0) if i is not None else float("nan")
        except ValueError:
            return float("nan")
    print(f"# {path}")
    print(f"{'kernel':52s} {'us':>8} {'rdMB':>7} {'wrMB':>7} {'dram%':>6} {'sm%':>5} {'tensor%':>7} {'tc%':>5} {'issue%':>6} {'occ%':>5} {'regs':>4}  top stalls (warps per issue)")
    for r in rd:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ki].replace("void pvrl::<unnamed>::", "").replace("pvrl::<unnamed>::", ""))[:52]
        st = sorted(((float(r[i].replace(",", "") or 0), h.split("stalled_")[1].split("_per")[0]) for i, h in stalls), reverse=True)[:3]
        print(f"{name:52s} {f(r,'dur_ns')/1e3:8.1f} {f(r,'rd')/1e6:7.1f} {f(r,'wr')/1e6:7.1f} {f(r,'dram%'):6.1f} {f(r,'sm%'):5.1f} "
              f"{f(r,'tensor%'):7.1f} {f(r,'tc%'):5.1f} {f(r,'issue%'):6.1f} {f(r,'occ%'):5.1f} {f(r,'regs'):4.0f}  "
              + ", ".join(f"{n} {v:.1f}" for v, n in st))


if __name__ == "__main__":
    main(sys.argv[1])
