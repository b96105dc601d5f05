#!/usr/bin/env python
"""Hottest SASS instructions of an `ncu --import-source on` report exported with `--page source --csv`:
   ncu -i x.ncu-rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [top]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    si = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in body)
    print(f"{len(body)} instructions, {tot} samples")
    agg = {}
    for r in body:
        for i, h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    print("stall reasons: " + ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    idx = sorted(range(len(body)), key=lambda k: -int(body[k][si]))[:top]
    for k in sorted(idx):
        r = body[k]
        st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        print(f"#{k:5d} {100 * int(r[si]) / tot:5.1f}%  {r[1].strip()[:70]:70s} " + " ".join(f"{n}:{v}" for v, n in st if v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
