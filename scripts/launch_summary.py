"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python scripts/launch_summary.py gpurun_out/launches.csv > profiles/launches_summary.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    name = re.sub(r"\(.*$", "", r[ki])
    name = name.replace("(anonymous namespace)", "<unnamed>")
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print(f"launches: {n}, total kernel time {tot:.1f} ms\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{name[:110]}` | {c} | {ms:.2f} | {100 * ms / tot:.1f}% |")
