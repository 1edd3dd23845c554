#!/usr/bin/env python
"""Summarise an ncu csv holding gpu__time_duration.sum + dram__bytes_{read,write}.sum per launch:
per kernel name -> launches, time, DRAM bytes, and (--json) the per-launch averages bench.py reports as
roofline.traffic.   python scripts/summarize_dram.py gpurun_out/ncu_dram_g5.csv [--json out.json]"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3,
        "ms": 1e6, "msecond": 1e6}


def main(path, json_out=None):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        per[name][row["Metric Name"]] += v
        ids[name].add(row["ID"])
    T = sum(d["gpu__time_duration.sum"] for d in per.values())
    print(f"# {path}: {sum(len(v) for v in ids.values())} launches, {T / 1e6:.3f} ms under ncu (serialised, cold cache)")
    print(f"{'kernel':36s} {'launches':>8s} {'ms':>9s} {'share':>6s} {'dram rd GB':>11s} {'dram wr GB':>11s} {'GB/s':>8s}")
    out = {}
    for k, d in sorted(per.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
        t, rd, wr = d["gpu__time_duration.sum"], d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        n = len(ids[k])
        print(f"{k:36s} {n:8d} {t / 1e6:9.3f} {100 * t / T:5.1f}% {rd / 1e9:11.3f} {wr / 1e9:11.3f} {(rd + wr) / t:8.1f}")
        out[k] = {"launches": n, "ms": t / 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr,
                  "dram_bytes_per_launch": (rd + wr) / n}
    if json_out:
        with open(json_out, "w") as f:
            json.dump({"source": path, "kernels": out}, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
