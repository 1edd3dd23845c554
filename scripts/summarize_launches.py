#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt"""
import collections
import csv
import sys


def main(path, top=20):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    tot, cnt, items = collections.defaultdict(float), collections.Counter(), []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v * 1e3 if unit in ("us", "usecond") else (v * 1e6 if unit in ("ms", "msecond") else v)
        tot[name] += ns
        cnt[name] += 1
        items.append((ns, name, row["Grid Size"], row["ID"]))
    T = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {T / 1e6:.3f} ms (cold-cache, serialised: compare shares)")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:34s} {cnt[k]:4d} launches {v / 1e6:8.3f} ms {100 * v / T:5.1f}%")
    print(f"# top {top} launches")
    for ns, name, grid, i in sorted(items, reverse=True)[:top]:
        print(f"{ns / 1e3:9.1f} us  {name:26s} grid {grid:14s} id {i}")


if __name__ == "__main__":
    main(sys.argv[1])
