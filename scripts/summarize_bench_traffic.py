#!/usr/bin/env python
"""ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) of a bench.py run ->
per-kernel DRAM traffic summary, stamped with the digest of the kernel sources it was captured with.
    python scripts/summarize_bench_traffic.py IN.csv OUT.json [groups_per_call] [workload]"""
import csv
import io
import json
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cap4d_b200 import build as B  # noqa: E402

src, dst = sys.argv[1], sys.argv[2]
gpc = int(sys.argv[3]) if len(sys.argv) > 3 else 5
workload = sys.argv[4] if len(sys.argv) > 4 else "single_ref"
with open(src, errors="replace") as fh:
    text = fh.read()
start = text.find('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:]))) if start >= 0 else []
per = defaultdict(lambda: defaultdict(float))
launch_ids = defaultdict(set)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
         "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
for r in rows:
    name = re.sub(r"<.*", "", r.get("Kernel Name", "")).split("(")[0].strip()
    name = name.split("::")[-1]
    metric, unit = r.get("Metric Name", ""), r.get("Metric Unit", "")
    try:
        val = float(r.get("Metric Value", "0").replace(",", ""))
    except ValueError:
        continue
    val *= scale.get(unit, 1.0)
    per[name][metric] += val
    launch_ids[name].add(r.get("ID"))
kernels = {}
tot_t = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values()) or 1.0
for name, v in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0)):
    n = max(1, len(launch_ids[name]))
    rd, wr, t = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0), v.get("gpu__time_duration.sum", 0.0)
    kernels[name] = {"launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd + wr) / n,
                     "time_us": t, "time_share": t / tot_t, "dram_gbs": (rd + wr) / t / 1e3 if t > 0 else None}
out = {"what": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over bench.py "
               "(scripts/ncu_bench_traffic.sh): launches of the first U-Net calls of the timed region, cold-cache and "
               "serialised (compare shares, not absolutes)",
       "source_digest": B.kernel_digest(), "groups_per_call": gpc, "workload": workload, "n_launches": sum(len(s) for s in launch_ids.values()),
       "library_kernels": sorted(k for k in kernels if k.startswith("at::") or "at::native" in k or "elementwise_kernel" in k or "vectorized_" in k or "cutlass" in k or "cudnn" in k),
       "kernels": kernels}
with open(dst, "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps({k: (round(v["time_share"], 3), v["launches"]) for k, v in list(kernels.items())[:12]}))
print("library (ATen) kernels in the capture:", out["library_kernels"])
