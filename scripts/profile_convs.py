"""Per-shape timing of every convolution launch of one training step (eager, each C-ABI conv call bracketed
by CUDA events and synchronised).  Prints, per (entry point, geometry): calls, total ms, TFLOP/s.
Usage: python scripts/profile_convs.py [batch] > gpurun_out/profile_convs.txt"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200.train_step import Trainer, default_args

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tr = Trainer(default_args(batch_size=B), device="cuda", seed=0)
X = torch.empty(B, 3, 256, 256, device="cuda").uniform_(-1, 1)
for it in (16, 1):
    tr.step(X, it)
torch.cuda.synchronize()

stats = collections.defaultdict(lambda: [0, 0.0])
orig = _lib.call
CONV = {"ideas_conv2d_forward": 6, "ideas_conv2d_dgrad": 6, "ideas_conv2d_wgrad": 5}


def timed_call(name, *args):
    if name not in CONV:
        return orig(name, *args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(name, *args)
    b.record()
    b.synchronize()
    n = CONV[name]
    geom = tuple(int(v) for v in args[n:n + 9])          # N H W C K kh kw stride pad
    scales = tuple(bool(v) for v in args[3:n])
    key = (name.replace("ideas_conv2d_", ""), geom, scales)
    stats[key][0] += 1
    stats[key][1] += a.elapsed_time(b)


_lib.call = timed_call
import ideas_b200.stylegan2.op.conv as C
C._lib.call = timed_call
tr.step(X, 2)
torch.cuda.synchronize()
_lib.call = orig
rows = []
for (name, g, sc), (cnt, ms) in stats.items():
    N, H, W, Cc, K, kh, kw, s, pad = g
    OH, OW = (H + 2 * pad - kh) // s + 1, (W + 2 * pad - kw) // s + 1
    flops = 2.0 * N * OH * OW * K * Cc * kh * kw * cnt
    rows.append((ms, name, g, cnt, flops / ms / 1e9 if ms > 0 else 0))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"conv launches of one step (no R1): {tot:.1f} ms total")
for ms, name, g, cnt, tf in rows[:80]:
    print(f"{ms:8.2f} ms {100 * ms / tot:5.1f}%  {cnt:3d} x  {tf:7.1f} TFLOP/s  {name:8s} N,H,W,C,K,kh,kw,s,p={g}")
