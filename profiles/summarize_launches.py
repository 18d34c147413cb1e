#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<.*", "", name)[:60]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e6:.2f} ms of kernel time (cold-cache, serialised)")
    print(f"{'total ms':>10} {'share':>6} {'n':>6} {'avg us':>10}  kernel")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v / 1e6:10.2f} {100 * v / T:5.1f}% {cnt[k]:6d} {v / cnt[k] / 1e3:10.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
