#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel name the launch
count, total / mean duration and share.  Usage: summarize_launches.py launches.csv [skip_launches]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        short = re.sub(r"\(.*", "", name).replace("comb::<unnamed>::", "").replace("void ", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("%-70s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %6d %12.1f %10.2f %6.1f%%" % (k[:70], n, ns / 1e3, ns / 1e3 / n, 100 * ns / tot))
    print("%-70s %6d %12.1f" % ("TOTAL", len(rows), tot / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
