"""Which tensors the small torch kernels of a training step (aten::mul / add / fill_ / copy_ / sum) work on:
torch.profiler with shapes, two eager steps, grouped by (op, input shapes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from ideas_b200.train_step import Trainer, default_args

B = 32
tr = Trainer(default_args(batch_size=B), device="cuda", seed=0, cuda_graphs=False)
X = torch.empty(B, 3, 256, 256, device="cuda").uniform_(-1, 1)
for it in (16, 1, 2):
    tr.step(X, it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.step(X, 3)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = {}
for e in rows:
    tot[e.key] = tot.get(e.key, 0) + e.self_device_time_total
print({k: round(v / 1e3, 2) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]})
for e in rows[:60]:
    print(f"{e.self_device_time_total / 1e3:8.2f} ms {e.count:5d} x {e.key:22s} {str(e.input_shapes)[:110]}")
