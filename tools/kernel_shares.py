#!/usr/bin/env python
"""Per-kernel sums of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total ms, share of the listed time.
Usage: python tools/kernel_shares.py launches.csv [divide_by]   (divide_by: the number of V-cycles the list covers)"""
import csv
import re
import sys


def main():
    path = sys.argv[1]
    div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(rows)
    tot = {}
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace(",", ";")
        ns = float(r["Metric Value"]) * (1e3 if r["Metric Unit"] in ("us", "usecond") else 1e6 if r["Metric Unit"] in ("ms", "msecond") else 1.0)
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ns)
    total = sum(t for _, t in tot.values())
    print("kernel,launches,ms,share_pct")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name},{n / div:g},{t / 1e6 / div:.4f},{100 * t / total:.2f}")
    print(f"total,{sum(n for n, _ in tot.values()) / div:g},{total / 1e6 / div:.4f},100.00")


if __name__ == "__main__":
    main()
