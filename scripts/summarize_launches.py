#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count,
mean device time and share of the listed time.  Usage: summarize_launches.py in.csv [title]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        agg.setdefault(name, []).append(float(r["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {title}\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    print("| kernel | launches | mean us | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |")


if __name__ == "__main__":
    main()
