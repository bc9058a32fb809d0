"""Kernel-time breakdown of one IDEAS training step (torch.profiler, CUDA activities).
Usage: python scripts/profile_step.py [batch] > gpurun_out/profile_step.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from ideas_b200.train_step import Trainer, default_args

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tr = Trainer(default_args(batch_size=B), device="cuda", seed=0)
X = torch.empty(B, 3, 256, 256, device="cuda").uniform_(-1, 1)
for it in (16, 1, 2):
    tr.step(X, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(X, 3)
    tr.step(X, 4)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
total = sum(e.self_device_time_total for e in rows)
print(f"total device time {total / 1e3:.1f} ms for 2 steps")
for e in rows[:70]:
    print(f"{e.self_device_time_total / 1e3:9.2f} ms {100 * e.self_device_time_total / total:5.1f}% {e.count:6d} x {e.self_device_time_total / max(e.count, 1):8.1f} us  {e.key[:110]}")
