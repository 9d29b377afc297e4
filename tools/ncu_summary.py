#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, avg, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^.*::", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} {'launches':>8s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:58]:58s} {v[0]:8d} {v[1]:11.1f} {v[1] / v[0]:9.2f} {100 * v[1] / tot:5.1f}%")
    print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {tot:11.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
