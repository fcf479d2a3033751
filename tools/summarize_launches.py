"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print("kernel,launches,total_us,share")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{us:.1f},{us / total:.4f}")
