"""Run one conv forward shape a few times (for ncu / timing).  Usage: prof_conv.py N C K H halo_mode [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ideas_b200 import _lib
from ideas_b200._tensor import ptr, stream_ptr

N, C, K, H, mode = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
_lib.call("ideas_set_option", b"halo", mode)
dev = torch.device("cuda")
x = torch.randn(N, H, H, C, device=dev)
wp = torch.randn(9, K, C, device=dev) / (C * 9) ** 0.5
d = torch.rand(N, K, device=dev) + 0.5
bias = torch.randn(K, device=dev)
y = torch.empty(N, H, H, K, device=dev)
for _ in range(iters):
    _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(bias), N, H, H, C, K, 3, 3, 1, 1, 1, 0.2,
              2 ** 0.5, 0, stream_ptr(x))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(bias), N, H, H, C, K, 3, 3, 1, 1, 1, 0.2,
              2 ** 0.5, 0, stream_ptr(x))
b.record()
torch.cuda.synchronize()
t = a.elapsed_time(b) / iters
print(f"N={N} C={C} K={K} H={H} halo={mode}: {t:.4f} ms, {2.0 * N * H * H * K * C * 9 / t / 1e9:.1f} TFLOP/s")
