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
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tmem.sum", "tmem inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "SM active cycles"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
]
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index("Kernel Name")]
    print(f"## {path}\nkernel: {name}")
    for key, label in KEYS:
        for h, u, v in zip(hdr, units, vals):
            if h == key or h.endswith("." + key):
                print(f"  {label:24s} {v} {u}")
                break
    print()
