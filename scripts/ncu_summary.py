"""Summarise an .ncu-rep (one kernel) into the handful of numbers DESIGN.md / bench.py quote.
Usage: python scripts/ncu_summary.py report.ncu-rep [more.ncu-rep ...]  (needs ncu on PATH)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram read % of peak"),
    ("dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram write % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe (realtime) %"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]
import json
import os

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
args = sys.argv[1:]
traffic_path = None
if args and args[0] == "--traffic":
    traffic_path, args = args[1], args[2:]
traffic = {}
if traffic_path and os.path.exists(traffic_path):
    traffic = json.load(open(traffic_path))
for path in args:
    if path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 5]
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")]
    print(f"## {path}\nkernel: {name}")
    got = {}
    for key, label in KEYS:
        cands = [(h, u, v) for h, u, v in zip(hdr, units, vals) if h == key] or \
                [(h, u, v) for h, u, v in zip(hdr, units, vals) if h.endswith("." + key) and v.strip()]
        if cands:
            h, u, v = cands[0]
            print(f"  {label:28s} {v} {u}")
            got[key] = (v, u)
    print()
    if traffic_path and "dram__bytes_read.sum" in got and "dram__bytes_write.sum" in got:
        tot = sum(float(got[k][0].replace(",", "")) * UNIT.get(got[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        case = os.path.basename(path).split(".")[0].split("_", 1)[1]
        traffic[case] = {"dram_bytes": int(tot), "kernel": name[:80],
                         "source": f"profiles/{os.path.basename(path).split('.')[0]}.raw.csv (ncu --set full, one launch)"}
if traffic_path:
    json.dump(traffic, open(traffic_path, "w"), indent=1)
