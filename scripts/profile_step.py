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
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
