#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the summed device time)."""
import collections
import csv
import re
import sys


def kname(full):
    s = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", full)
    m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(<|\(|$)", s.replace("void ", "").split("::")[-1] if "<" not in s else
                  re.sub(r"<.*", "", s.replace("void ", "")).split("::")[-1] + "<")
    base = m.group(1) if m else s[:50]
    t = re.search(r"<(.*)>", s)
    if base.startswith(("gemm_bf16", "attn_", "layernorm", "gather_cast", "colsum")) and t:
        base += "<" + t.group(1)[:40] + ">"
    return base


def main(path, skip=0, steps=1, exclude=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for i, row in enumerate(csv.DictReader(lines)):
        if i < skip:
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row.get("Metric Unit", "us")
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        n = kname(row["Kernel Name"])
        if exclude and re.search(exclude, n):
            continue
        agg[n][0] += 1
        agg[n][1] += v
        total += v
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms summed device time"
          + (f" over {steps} identical steps -> {total / 1e3 / steps:.2f} ms and {sum(a[0] for a in agg.values()) // steps} launches per step"
             if steps > 1 else ""))
    print(f"{'share':>7} {'ms/step':>9} {'n/step':>6} {'avg_us':>9}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"{t / total * 100:6.2f}% {t / 1e3 / steps:9.3f} {n / steps:6.0f} {t / n:9.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 1,
         sys.argv[4] if len(sys.argv) > 4 else None)
