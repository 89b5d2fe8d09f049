#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  Usage: launch_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"launches {len(rows) - 1}  total {tot / 1e6:.3f} ms (per-launch times are cold-cache, serialised: compare shares)")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / tot * 100:6.2f}%  n={a[0]:4d}  sum={a[1] / 1e6:9.3f} ms  avg={a[1] / a[0] / 1e3:9.1f} us  {n[:80]}")


if __name__ == "__main__":
    main(sys.argv[1])
