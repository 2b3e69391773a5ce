#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list: per kernel name, launches, mean duration, DRAM read/write per launch."""
import csv
import sys
from collections import defaultdict


def summarise(path):
    rows = defaultdict(lambda: defaultdict(list))
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        rows[name][r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
    print(f"== {path}")
    tot = 0.0
    out = []
    for name, m in rows.items():
        d = m.get("gpu__time_duration.sum", [0])
        rd = m.get("dram__bytes_read.sum", [0])
        wr = m.get("dram__bytes_write.sum", [0])
        mean = sum(d) / len(d)
        tot += mean
        out.append((name, len(d), mean, sum(rd) / len(rd), sum(wr) / len(wr)))
    for name, n, mean, rd, wr in out:
        print(f"{name:60s} n={n:3d} {mean / 1e3:9.2f} us {100 * mean / tot:5.1f}%  dram rd {rd / 1e6:9.3f} MB wr {wr / 1e6:9.3f} MB")
    print(f"{'sum of means':60s}       {tot / 1e3:9.2f} us")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        try:
            summarise(p)
        except Exception as ex:  # a missing or empty capture must not fail the whole visit
            print(f"== {p}: {ex}")
